"""Small bring-up run of one kernel variant against the exact scan: python scripts/pair_debug.py tmem 2 [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_generation_b200 import ResidualQuantizer  # noqa: E402

kernel, cluster = sys.argv[1], int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
nq, K, d = 2, 256, 128
torch.manual_seed(0)
m = ResidualQuantizer(nq, d, "ema", K, kernel=kernel, cluster=cluster).cuda().eval()
x = torch.randn(N, d, device="cuda")
with torch.no_grad():
    m.algo = "exact_scan"
    _, ie, _ = m(x)
    torch.cuda.synchronize()
    print("exact scan done", flush=True)
    m.algo = "tensor"
    _, it, _ = m(x)
    torch.cuda.synchronize()
print("equal:", bool(torch.equal(ie, it)), "mismatching frames:", int((ie != it).any(-1).sum()), flush=True)
