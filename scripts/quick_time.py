"""Quick timing of the encode kernels on the bench shapes: python scripts/quick_time.py [kernel/cluster ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_generation_b200 import ResidualQuantizer  # noqa: E402


def make(nq, K, d, kernel, cluster=0, algo="tensor"):
    torch.manual_seed(0)
    m = ResidualQuantizer(nq, d, "ema", K, kernel=kernel, cluster=cluster, algo=algo, vq_cutoff_freq=0.0, use_som=False)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.7 ** q)
        m.ema_sum.copy_(m.codebooks)
    return m.cuda().eval()


variants = [v.split("/") for v in sys.argv[1:]] or [["frame", "1"], ["tmem", "2"], ["generic", "0"]]
for nq, K, d, N, update in [(8, 1024, 128, 1 << 20, False), (12, 1024, 256, 1 << 20, False), (12, 1024, 256, 1 << 20, True),
                            (8, 1024, 128, 1 << 20, True)]:
    x = torch.randn(N, d, device="cuda")
    ref = None
    for kernel, cluster in variants:
        if kernel == "tmem" and d > 128:
            continue
        m = make(nq, K, d, kernel, int(cluster))
        m.train(update)
        evs = []
        m.kernel_events = evs
        for _ in range(13):
            with torch.no_grad():
                out = m(x, None, update_codebook=update)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs[3:])[len(evs[3:]) // 2]
        tf = N * nq * 2 * K * d / (ms * 1e-3) / 1e12
        if not update:
            if ref is None:
                ref = out[1]
            same = bool(torch.equal(ref, out[1]))
        else:
            same = "-"
        print(f"nq={nq} K={K} d={d} update={update} {kernel}/{cluster}: kernel {ms:.3f} ms  {N / ms / 1e3:.1f} M frames/s  "
              f"{tf:.0f} TFLOP/s ({tf / 1635.7:.3f} of measured peak)  same_idx_as_first={same}", flush=True)
