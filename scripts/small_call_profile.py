"""Cycle counters of the generic encode kernel on the reference's small call shapes (RVQ_FLAG_COUNTERS).
    python scripts/small_call_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_generation_b200 import ResidualQuantizer

for (B, L, d, nq, K) in [(1, 136, 512, 10, 512), (4, 150, 512, 10, 512), (8, 500, 512, 8, 1024), (1, 136, 256, 10, 512)]:
    torch.manual_seed(0)
    q = ResidualQuantizer(nq, d, "ema", K, use_som=False, vq_cutoff_freq=0.0).cuda().eval()
    x = torch.randn(B, d, L, device="cuda").permute(0, 2, 1)     # the reference's "b c l -> b l c" view (vae.py:313)
    q.kernel_events = evs = []
    with torch.no_grad():
        for _ in range(20):
            q(x)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs[5:])[len(evs[5:]) // 2]
    q.kernel_events = None
    q.counters = True
    with torch.no_grad():
        q(x)
    torch.cuda.synchronize()
    c = q.read_counters()
    n = max(c[5], 1)
    print(f"B={B} L={L} d={d} nq={nq} K={K}: encode kernel {ms * 1e3:.1f} us = {ms * 1e-3 * 1.965e9 / nq:.0f} cycles per stage; "
          f"tile-stages={c[5]}")
    print(f"   per tile-stage: scan={c[0] / n:.0f} (of which waiting for the accumulator {c[11] / n:.0f}; + wait for the operand "
          f"{c[1] / n:.0f})  update={c[2] / n:.0f} (+ wait for the scan {c[7] / n:.0f}): score={c[8] / n:.0f} apply={c[9] / n:.0f} "
          f"exact-scan={c[3] / n:.0f}  exact-scan rows={c[4] / n:.3f}")
