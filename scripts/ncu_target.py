"""Small fixed workload for ncu captures: python scripts/ncu_target.py <workload> [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200 import ResidualQuantizer

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
nq, K, d = wl["nq"], wl["K"], wl["d"]
q = ResidualQuantizer(nq, d, "ema", K)
with torch.no_grad():
    q.codebooks.copy_(bench.synth_codebooks(nq, K, d))
q = q.cuda().eval()
x = torch.randn(N, d, device="cuda")
for _ in range(3):
    with torch.no_grad():
        q(x)
torch.cuda.synchronize()
print("ok", name, N)
