"""Small fixed workload for ncu captures: python scripts/ncu_target.py <workload> [frames] [update]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200 import ResidualQuantizer

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
nq, K, d = int(os.environ.get("NCU_NQ", wl["nq"])), wl["K"], wl["d"]
update = len(sys.argv) > 3 and sys.argv[3] == "update"
q = ResidualQuantizer(nq, d, "ema", K, use_som=bool(wl.get("som", False)), vq_cutoff_freq=float(wl.get("cutoff", 0.0)))
with torch.no_grad():
    q.codebooks.copy_(bench.synth_codebooks(nq, K, d))
    q.ema_sum.copy_(q.codebooks)
q = q.cuda().train(update)
x = torch.randn(N, d, device="cuda")
for _ in range(3):
    with torch.no_grad():
        q(x, None, update_codebook=update)
torch.cuda.synchronize()
print("ok", name, N)
