"""Encode kernel time on row-major frames [N, d] against the reference's storage (a "b c l -> b l c" view of (B, d, L),
vae.py:313) at the bench sizes.   python scripts/layout_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_generation_b200 import ResidualQuantizer

for nq, K, d, N in [(8, 1024, 128, 1 << 20), (12, 1024, 256, 1 << 20), (8, 1024, 512, 1 << 18)]:
    torch.manual_seed(0)
    m = ResidualQuantizer(nq, d, "ema", K, vq_cutoff_freq=0.0, use_som=False)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.7 ** q)
    m = m.cuda().eval()
    B = 16
    xb = torch.randn(B, d, N // B, device="cuda")
    inputs = {"rows [N, d]": xb.permute(0, 2, 1).reshape(N, d).contiguous(), "(B, d, L) view": xb.permute(0, 2, 1)}
    ref = None
    for name, x in inputs.items():
        m.kernel_events = evs = []
        with torch.no_grad():
            for _ in range(13):
                out = m(x)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs[3:])[len(evs[3:]) // 2]
        idx = out[1].reshape(N, nq)
        same = True if ref is None else bool(torch.equal(ref, idx))
        ref = idx if ref is None else ref
        print(f"nq={nq} K={K} d={d} N={N} {name}: kernel {ms:.3f} ms  {N / ms / 1e3:.1f} M frames/s  same indices={same}", flush=True)
