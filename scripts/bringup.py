"""GPU bring-up diagnostics (not a test): checks each layer of the tensor path against torch.

    python scripts/bringup.py [d] [K]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_generation_b200 import _lib
from audio_generation_b200.quantizer import ResidualQuantizer, _ptr, _stream
from oracle import rvq_oracle as O


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    nq = 4
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    lib = _lib.load()
    print("abi", lib.rvq_version(), "supported", lib.rvq_device_supported(0), torch.cuda.get_device_name(0))
    m = ResidualQuantizer(nq, d, "ema", K).to(dev).eval()
    m.codebooks.mul_(0.5)
    op, nrm, meta = m._prepared()
    torch.cuda.synchronize()
    Kpad = (K + 255) // 256 * 256
    meta_c = meta.reshape(nq, 8).cpu()
    print("meta[0]", meta_c[0].tolist())
    cb = m.codebooks.detach()
    sb = meta_c[0, 0].item()
    ref_op = (cb[0] * (-2.0 * sb)).half()
    print("prep op max diff", (op.reshape(nq, Kpad, d)[0, :K].float() - ref_op.float()).abs().max().item())
    ref_n = (cb[0] * cb[0]).sum(1) * sb * sb
    print("prep norm rel diff", ((nrm[:nq * Kpad].reshape(nq, Kpad)[0, :K] - ref_n).abs() / ref_n).max().item())

    # ---- single-stage filter scores
    x = torch.randn(128, d, device=dev)
    scores = torch.full((128, Kpad), float("nan"), device=dev)
    rs = torch.zeros(128, device=dev)
    rc = lib.rvq_debug_stage_scores(_ptr(x), d, K, 0, _ptr(op), _ptr(nrm), _ptr(meta), _ptr(scores), _ptr(rs), _stream())
    print("debug rc", rc, lib.rvq_last_error())
    torch.cuda.synchronize()
    a_h = (x * rs[:, None]).half().float()
    b_h = op.reshape(nq, Kpad, d)[0].float()
    na = rs / sb
    ref = a_h.double() @ b_h.double().t() + (na[:, None] * nrm[:nq * Kpad].reshape(nq, Kpad)[0][None, :]).double()
    err = (scores.double() - ref).abs()
    print("scores nan:", torch.isnan(scores).sum().item(), "max abs err", err[:, :K].max().item(),
          "ref scale", ref[:, :K].abs().max().item())
    bad = (err[:, :K] > 1e-3 * ref[:, :K].abs().max()).nonzero()
    print("bad entries", bad.shape[0], bad[:8].tolist())
    print("argmin agree (approx vs fp16-ref)", (scores[:, :K].argmin(1) == ref[:, :K].argmin(1)).float().mean().item())

    # ---- full encode vs oracle (small)
    for N in (128, 1000, 20000):
        xx = torch.randn(N, d, device=dev)
        for algo in ("exact_scan", "tensor"):
            m.algo = algo
            t0 = time.time()
            xq, idx, commit = m(xx)
            torch.cuda.synchronize()
            dt = time.time() - t0
            cbs = [m.codebooks[q].cpu() for q in range(nq)]
            ri, rxq, rr, rc_ = O.rvq_encode_ref(xx.cpu(), cbs)
            mism = (ri != idx.cpu()).sum().item()
            adj = O.adjudicate_indices(xx.cpu(), cbs, idx.cpu())
            print(f"N={N} algo={algo} t={dt*1e3:.2f}ms idx mismatches={mism} adjudicated={adj} "
                  f"xq err={(xq.cpu()-rxq).abs().max().item():.3e} commit={commit.item():.6f} ref={sum(rc_):.6f}")
        m.algo = "tensor"
        xq_t, idx_t, _ = m(xx)
        m.algo = "exact_scan"
        xq_e, idx_e, _ = m(xx)
        print(f"N={N} tensor vs exact_scan index disagreements: {(idx_t != idx_e).sum().item()}")


if __name__ == "__main__":
    main()
