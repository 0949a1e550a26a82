"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes) by the drop-in module,
against the CPU oracle (oracle/rvq_oracle.py) on the same seeded inputs.

Bars (BASELINE.json north star):
  * code indices: equal to the oracle, except frames where the two candidate scores differ in fp64 by less
    than eps_tie = 8*d*2^-24*(||r||*||c||max + ||c||max^2) (fp32 dot-product reordering noise); later stages
    of such a frame are compared teacher-forced (oracle re-run on the kernel's own prefix);
  * the tensor-core path and the exact-scan path share one exact scoring routine, so THEY must agree bit for bit;
  * xq / residual: atol 1e-5*max|x| (+ exact-arithmetic identity xq = sum of code vectors);
  * EMA statistics / codebooks: rtol 1e-5 (+ atol 1e-6*count) -- fp32 atomics reorder the sums;
  * commit loss: rtol 1e-5.
"""
import os

import pytest
import torch

from oracle import rvq_oracle as O

pytestmark = pytest.mark.gpu

cuda = torch.cuda.is_available()


def make(nq, K, d, algo="tensor", seed=0, scale_decay=0.7, cls="ema"):
    from audio_generation_b200 import ResidualQuantizer
    torch.manual_seed(seed)
    m = ResidualQuantizer(nq, d, cls, K, algo=algo)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(scale_decay ** q)
        m.ema_sum.copy_(m.codebooks)
    return m.cuda()


def cbs_of(m):
    return [m.codebooks[q, :m.codebook_sizes[q]].detach().cpu() for q in range(m.num_quantizers)]


def check_against_oracle(m, x, nq=None, max_mismatch_frac=2e-4):
    m.eval()
    with torch.no_grad():
        xq, idx, commit = m(x, nq)
    torch.cuda.synchronize()
    d = m.dim
    x2 = x.detach().reshape(-1, d).cpu()
    cbs = cbs_of(m)
    nqu = idx.shape[-1]
    ri, rxq, rr, rc = O.rvq_encode_ref(x2, cbs, nqu)
    i2 = idx.reshape(-1, nqu).cpu()
    adj = O.adjudicate_indices(x2, cbs, i2)
    assert adj["n_illegal"] == 0, adj
    assert adj["n_mismatch"] <= max(2, max_mismatch_frac * i2.numel()), adj
    same = (i2 == ri).all(dim=1)
    tol = 1e-5 * float(x2.abs().max())
    xq2 = xq.detach().reshape(-1, d).cpu()
    assert (xq2[same] - rxq[same]).abs().max() <= tol
    # exact identity on every frame: xq is the sum of the selected code vectors
    deq = sum(cbs[q][i2[:, q]] for q in range(nqu))
    assert (xq2 - deq).abs().max() <= tol
    if same.all():
        assert abs(float(commit) - sum(rc)) <= 1e-5 * abs(sum(rc)) + 1e-12
    return idx, adj


@pytest.mark.parametrize("nq,K,d,N", [(4, 1024, 128, 5000), (3, 512, 512, 1000), (3, 1024, 256, 3000),
                                      (2, 256, 64, 777), (2, 300, 128, 1029), (2, 4096, 512, 300)])
def test_encode_matches_oracle(nq, K, d, N):
    m = make(nq, K, d)
    x = torch.randn(N, d, device="cuda")
    check_against_oracle(m, x)


@pytest.mark.parametrize("algo", ["exact_scan"])
def test_exact_scan_matches_oracle(algo):
    m = make(3, 512, 128, algo=algo)
    x = torch.randn(3000, 128, device="cuda")
    check_against_oracle(m, x)


@pytest.mark.parametrize("nq,K,d,N", [(8, 1024, 128, 1 << 16), (4, 1024, 256, 1 << 14), (2, 512, 512, 1 << 13)])
def test_tensor_path_equals_exact_scan_bitwise(nq, K, d, N):
    """Same exact scorer in both paths => the certified filter must pick identical codes everywhere."""
    m = make(nq, K, d)
    x = torch.randn(N, d, device="cuda")
    m.eval()
    with torch.no_grad():
        m.algo = "tensor"
        xq_t, idx_t, c_t = m(x)
        m.algo = "exact_scan"
        xq_e, idx_e, c_e = m(x)
    assert torch.equal(idx_t, idx_e)
    assert torch.equal(xq_t, xq_e)
    assert abs(float(c_t) - float(c_e)) <= 1e-6 * abs(float(c_e))


def test_reference_layout_view_and_partial_stages():
    """The reference passes a (B, L, d) VIEW of a (B, d, L) tensor (vae.py:313) and a per-call stage count."""
    m = make(5, 512, 512)
    xc = torch.randn(3, 512, 150, device="cuda")          # (B, C, L) as the encoder produces it
    xv = xc.permute(0, 2, 1)                               # b c l -> b l c view, strides (C*L, 1, L)
    assert not xv.is_contiguous()
    m.eval()
    with torch.no_grad():
        xq_v, idx_v, c_v = m(xv, 3)
        xq_c, idx_c, c_c = m(xv.contiguous(), 3)
    assert idx_v.shape == (3, 150, 3) and idx_v.dtype == torch.int64
    assert xq_v.shape == xv.shape
    assert torch.equal(idx_v, idx_c) and torch.equal(xq_v, xq_c)
    assert xq_v.permute(0, 2, 1).is_contiguous()           # decoder-side rearrange is a free view
    check_against_oracle(m, xv, 3)
    import numpy as np
    with torch.no_grad():
        _, idx_n, _ = m(xv, np.int64(2))                   # training.py:294 passes a NumPy int
    assert idx_n.shape[-1] == 2 and torch.equal(idx_n, idx_v[..., :2])


def test_edge_inputs():
    m = make(3, 512, 128)
    m.eval()
    with torch.no_grad():
        # empty
        xq, idx, c = m(torch.empty(0, 128, device="cuda"))
        assert xq.shape == (0, 128) and idx.shape == (0, 3)
        # single frame, zero frame, frame equal to a code, huge and tiny magnitudes
        x = torch.randn(64, 128, device="cuda")
        x[0] = 0
        x[1] = m.codebooks[0, 17]
        x[2] *= 1e6
        x[3] *= 1e-6
        x[4] = m.codebooks[0, 5] + m.codebooks[1, 9]
    check_against_oracle(m, x)
    check_against_oracle(m, x[:1])
    xq, idx, _ = m(x)
    assert int(idx[1, 0]) == 17


def test_duplicate_and_degenerate_codebooks():
    """Exact ties: duplicated codes must resolve to the LOWEST index (torch argmin on CPU); an all-equal
    codebook exercises the exact-scan fallback for every frame."""
    m = make(2, 512, 128)
    with torch.no_grad():
        m.codebooks[0, 300] = m.codebooks[0, 40]
        m.codebooks[0, 41] = m.codebooks[0, 40]
        m.codebooks[1, :] = m.codebooks[1, 0]
    m.invalidate()
    x = torch.randn(2000, 128, device="cuda")
    with torch.no_grad():
        x[:50] = m.codebooks[0, 40] + 0.01 * torch.randn(50, 128, device="cuda")
    idx, adj = check_against_oracle(m, x)
    i2 = idx.reshape(-1, 2)
    assert (i2[:50, 0] == 40).all()
    assert (i2[:, 1] == 0).all()
    assert adj["n_mismatch"] == 0


def test_ragged_codebook_sizes():
    from audio_generation_b200 import ResidualQuantizer
    torch.manual_seed(1)
    m = ResidualQuantizer(3, 128, "ema", [512, 300, 64]).cuda()
    x = torch.randn(1500, 128, device="cuda")
    idx, _ = check_against_oracle(m, x)
    i2 = idx.reshape(-1, 3)
    assert int(i2[:, 1].max()) < 300 and int(i2[:, 2].max()) < 64


@pytest.mark.parametrize("algo,d", [("tensor", 128), ("exact_scan", 128), ("tensor", 256), ("tensor", 512),
                                    ("tensor", 192), ("tensor", 64)])
def test_ema_update_matches_oracle(algo, d):
    """The north-star update (counts + summed vectors -> EMA refresh), SOM spreading and re-seeding switched off;
    tests/test_gpu_codebook_maintenance.py covers them.  d = 128 / 64: TMEM-resident kernel; 192 / 256 / 512: generic
    kernel.  Indices are adjudicated against the oracle (fp32 near-ties are legal); the statistics and the refreshed
    state are then checked teacher-forced, i.e. against the oracle's update computed from the GPU's own indices, so
    that one legal near-tie does not show up as a count difference of 0.01."""
    nq, K, N = 3, 512, 20000
    m = make(nq, K, d, algo=algo)
    m.use_som, m.vq_cutoff_freq = False, 0.0
    m.train()
    for step in range(3):
        cbs = cbs_of(m)
        cnt0, sum0 = m.ema_count.cpu().clone(), m.ema_sum.cpu().clone()
        x = torch.randn(N, d, device="cuda")
        with torch.no_grad():
            _, idx, c = m(x, None, update_codebook=True)
        torch.cuda.synchronize()
        xc, ic = x.cpu(), idx.cpu()
        adj = O.adjudicate_indices(xc, cbs, ic)
        assert adj["n_illegal"] == 0 and adj["n_mismatch"] <= 12, adj
        res = O.stage_residuals_from_indices(xc, cbs, ic)
        flat = m.last_stats.cpu()
        for q in range(nq):
            cnt, sm = O.ema_stats_ref(res[q], ic[:, q], K)
            assert torch.equal(flat[nq * K * d + q * K: nq * K * d + (q + 1) * K], cnt)              # counts are exact
            # sums: fp32 atomics add in arbitrary order -> absolute error ~ count * ulp(max |r|) per code
            tol = (2e-6 * cnt * float(res[q].abs().max()) + 1e-5)[:, None]
            gsum = flat[q * K * d:(q + 1) * K * d].reshape(K, d)
            assert bool(((gsum - sm).abs() <= tol).all()), float((gsum - sm).abs().max())
            ncb, nc, ns = O.ema_finalize_ref(cbs[q], cnt0[q], sum0[q], cnt, sm)
            assert torch.allclose(m.ema_count[q].cpu(), nc, rtol=1e-6, atol=1e-6)
            assert bool(((m.ema_sum[q].cpu() - ns).abs() <= 0.01 * tol + 1e-6 * ns.abs()).all())
            assert torch.allclose(m.codebooks[q].cpu(), ncb, rtol=1e-4, atol=1e-4)
        # commit loss of the call = sum over stages of mean squared residual after the stage
        rc = sum(float((res[q + 1].double() ** 2).mean()) for q in range(nq))
        assert abs(float(c) - rc) <= 1e-5 * rc


def test_eval_mode_does_not_update():
    m = make(2, 256, 64)
    before = m.codebooks.clone()
    m.eval()
    with torch.no_grad():
        m(torch.randn(500, 64, device="cuda"), None, update_codebook=True)
    assert torch.equal(before, m.codebooks)


@pytest.mark.parametrize("strided", [False, True])
@pytest.mark.parametrize("cls", ["ema", "base"])
def test_autograd_matches_oracle(cls, strided):
    """rvq_backward (straight-through + commit loss [+ codebook loss for "base"]) against torch autograd on the oracle;
    strided = the reference's (B, L, d) view of a (B, d, L) encoder output (vae.py:313)."""
    nq, K, d = 3, 256, 64
    m = make(nq, K, d, cls=cls)
    ref = O.ResidualQuantizerRef(nq, d, cls, K)
    with torch.no_grad():
        ref.codebooks.copy_(m.codebooks.detach().cpu())
    if strided:
        leaf = torch.randn(4, d, 50, device="cuda", requires_grad=True)
        x = leaf.permute(0, 2, 1)
        rleaf = leaf.detach().cpu().requires_grad_(True)
        xr = rleaf.permute(0, 2, 1)
    else:
        leaf = x = torch.randn(4, 50, d, device="cuda", requires_grad=True)
        rleaf = xr = x.detach().cpu().requires_grad_(True)
    w = torch.randn(4, 50, d, device="cuda")
    out, idx, commit = m(x)
    (out * w).sum().add(3.0 * commit).backward()
    ro, ridx, rcommit = ref(xr)
    (ro * w.cpu()).sum().add(3.0 * rcommit).backward()
    assert torch.equal(idx.cpu(), ridx)
    assert torch.allclose(commit.cpu(), rcommit, rtol=1e-5)
    assert torch.allclose(out.detach().cpu(), ro.detach(), atol=1e-5)
    assert torch.allclose(leaf.grad.cpu(), rleaf.grad, rtol=1e-4, atol=1e-6)
    if cls == "base":
        assert torch.allclose(m.codebooks.grad.cpu(), ref.codebooks.grad, rtol=1e-4, atol=1e-6)
    # commit loss alone (no gradient arriving through x_quantized), partial stages
    leaf.grad = None
    rleaf.grad = None
    _, _, c2 = m(x, 2)
    c2.backward()
    _, _, rc2 = ref(xr, 2)
    rc2.backward()
    assert torch.allclose(leaf.grad.cpu(), rleaf.grad, rtol=1e-4, atol=1e-7)


def test_dequantize_and_stage_api():
    m = make(4, 512, 128)
    x = torch.randn(2, 70, 128, device="cuda")
    m.eval()
    with torch.no_grad():
        xq, idx, _ = m(x)
    deq = m.dequantize(idx)
    assert torch.allclose(deq, xq, atol=1e-5)
    one = m.quantizers[2].dequantize(idx[:1, :, 2])          # (1, L) long -> (1, L, d)  (vae.py:333)
    assert one.shape == (1, 70, 128)
    assert torch.equal(one, m.codebooks[2][idx[:1, :, 2]])
    assert m.quantizers[0].som.height * m.quantizers[0].som.width == 512
    assert len(m.get_stale_clusters()) == 4
    m.update_cutoff(ratio=0.95)


def test_large_shape_properties():
    """BASELINE configs[1] at full size (1M frames): size-independent properties instead of the oracle."""
    nq, K, d, N = 8, 1024, 128, 1 << 20
    m = make(nq, K, d)
    m.eval()
    x = torch.randn(N, d, device="cuda")
    with torch.no_grad():
        xq, idx, commit = m(x)
        assert int(idx.min()) >= 0 and int(idx.max()) < K
        # decode(encode(x)) == xq ; residual norm decreases monotonically with more stages
        deq = m.dequantize(idx)
        assert (deq - xq).abs().max() <= 1e-5 * float(x.abs().max())
        prev = None
        for n in (1, 4, 8):
            xqn, idxn, cn = m(x, n)
            assert torch.equal(idxn, idx[:, :n])             # prefix property of the residual chain
            err = float(((x - xqn) ** 2).mean())
            assert prev is None or err < prev
            prev = err
        # idempotence on a sample: the oracle agrees on the first 65536 frames
    check_against_oracle(m, x[:1 << 16])


def test_host_encoder_roundtrip():
    from audio_generation_b200.quantizer import HostEncoder
    m = make(4, 512, 128)
    m.eval()
    xh = torch.randn(50000, 128).pin_memory()
    he = HostEncoder(m, chunk_frames=1 << 13)
    ih = he.encode(xh)
    torch.cuda.synchronize()
    with torch.no_grad():
        _, idx, _ = m(xh.cuda())
    assert torch.equal(ih, idx.cpu())


def test_refuses_cpu_tensor():
    from audio_generation_b200._lib import RVQError
    m = make(2, 256, 64)
    with pytest.raises(RVQError):
        m(torch.randn(10, 64))


def test_filter_scores_match_fp16_reference():
    """Bring-up hook: the tensor-core filter alone (TMA -> tcgen05 -> TMEM -> epilogue add) reproduces
    (2^a r)_fp16 . (-2 2^b C)_fp16^T + 2^(a-b) * 2^(2b)||c||^2 to fp32 accumulation accuracy."""
    import ctypes as C
    from audio_generation_b200 import _lib
    from audio_generation_b200.quantizer import _ptr, _stream
    lib = _lib.load()
    for (K, d) in [(1024, 128), (512, 256), (300, 512), (256, 64)]:
        m = make(2, K, d)
        op, nrm, meta = m._prepared()
        Kpad = (K + 255) // 256 * 256
        x = torch.randn(128, d, device="cuda") * 3.0
        scores = torch.full((128, Kpad), float("nan"), device="cuda")
        rs = torch.zeros(128, device="cuda")
        _lib.check(lib.rvq_debug_stage_scores(_ptr(x), d, K, 1, _ptr(op), _ptr(nrm), _ptr(meta), _ptr(scores), _ptr(rs),
                                              _stream()), "rvq_debug_stage_scores")
        torch.cuda.synchronize()
        sb = meta.reshape(2, 8)[1, 0]
        a_h = (x * rs[:, None]).half().double()
        b_h = op.reshape(2, Kpad, d)[1].double()
        ref = a_h @ b_h.t() + ((rs / sb)[:, None] * nrm[:2 * Kpad].reshape(2, Kpad)[1][None, :]).double()
        # codes above the stage's norm cap: the kernel's scores are optimistic by rs_row * xc_k in the 256-code chunks
        # that hold one (NormLayout, csrc/common.cuh; this entry point runs the generic kernel)
        xc = nrm[9 * 2 * Kpad:10 * 2 * Kpad].reshape(2, Kpad)[1].double()
        f0 = 10 * 2 * Kpad + 2 * Kpad // 4          # floats before the flags: norm, slices (8x), xc, byte table
        flag = nrm[f0:f0 + 2 * (Kpad // 256)].view(torch.int32).reshape(2, Kpad // 256)[1]
        rs_row = (x.double().pow(2).sum(1).sqrt() * 1.00002 * rs.double())
        ref = ref - rs_row[:, None] * (xc * flag.repeat_interleave(256).bool())[None, :]
        err = (scores.double() - ref).abs()[:, :K].max()
        assert not torch.isnan(scores[:, :K]).any()
        assert err <= 2e-6 * ref[:, :K].abs().max(), (K, d, float(err))
        assert (scores[:, :K].argmin(1) == ref[:, :K].argmin(1)).float().mean() > 0.99


def test_unsupported_shapes_are_refused():
    from audio_generation_b200 import ResidualQuantizer
    from audio_generation_b200._lib import RVQError
    with pytest.raises(RVQError):
        ResidualQuantizer(2, 100, "ema", 64).cuda()(torch.randn(8, 100, device="cuda"))       # d % 64 != 0
    with pytest.raises(RVQError):
        ResidualQuantizer(1, 64, "ema", 8448).cuda()(torch.randn(8, 64, device="cuda"))       # K > 8192


@pytest.mark.parametrize("nq,K", [(8, 1024), (10, 512), (32, 4096), (5, 300), (3, 2)])
def test_wire_format_matches_oracle_and_round_trips(nq, K):
    """Packed codes: bit-exact against the oracle's packing, pack -> unpack is the identity, and
    decode_packed(pack(idx)) equals the kernel's own xq."""
    m = make(nq, K, 64)
    m.eval()
    x = torch.randn(3, 333, 64, device="cuda")
    with torch.no_grad():
        xq, idx, _ = m(x)
        packed = m.pack_indices(idx)
        back = m.unpack_indices(packed)
        dec = m.decode_packed(packed)
    bits = m.code_bits
    assert packed.dtype == torch.uint8 and packed.shape == (3, 333, (nq * bits + 7) // 8)
    ref = O.pack_indices_ref(idx.reshape(-1, nq).cpu().numpy(), bits)
    assert (packed.reshape(-1, packed.shape[-1]).cpu().numpy() == ref).all()
    assert torch.equal(back, idx)
    assert (dec - xq).abs().max() <= 1e-5 * float(x.abs().max())
    # partial stage count
    with torch.no_grad():
        p3 = m.pack_indices(idx[..., :min(3, nq)])
        assert torch.equal(m.unpack_indices(p3, min(3, nq)), idx[..., :min(3, nq)])
