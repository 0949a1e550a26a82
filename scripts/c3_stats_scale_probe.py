"""Does the C3 encode kernel's time depend on how many frames stand behind one EMA update?  One GPU, no collective:
the statistics of the local 1M frames are multiplied by S before the EMA refresh, which is what the codebooks see
when S ranks with statistically equivalent shards all-reduce their statistics.  python scripts/c3_stats_scale_probe.py"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200 import ResidualQuantizer


class A:
    algo, kernel = "tensor", "auto"


dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["c3"])
x = torch.randn(wl["frames"], wl["d"], device=dev)
for S in (1, 8, 1, 2):
    q = bench.build_quantizer(wl, dev, A)
    orig = q._update_codebooks

    def scaled(x3, N, L, sb, sl, sd, nq, idx, flat, ssum, scnt, rep, _o=orig, _S=S):
        if _S != 1:
            flat[: ssum.numel() + scnt.numel()].mul_(float(_S))
        return _o(x3, N, L, sb, sl, sd, nq, idx, flat, ssum, scnt, rep)

    q._update_codebooks = scaled
    q.counters = True
    for _ in range(150):
        with torch.no_grad():
            q(x, None, update_codebook=True)
    torch.cuda.synchronize()
    q.kernel_events = []
    for _ in range(20):
        with torch.no_grad():
            q(x, None, update_codebook=True)
    torch.cuda.synchronize()
    k = statistics.mean(a.elapsed_time(b) for a, b in q.kernel_events)
    c = q.read_counters()
    print(f"statistics x {S}: encode kernel {k:.3f} ms in the steady state (150 warm-up updates); counters {c[:8]}", flush=True)
