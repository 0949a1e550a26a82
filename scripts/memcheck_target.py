import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_generation_b200 import ResidualQuantizer
nq, K, d, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
algo = sys.argv[5] if len(sys.argv) > 5 else "tensor"
scale = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
torch.manual_seed(0)
m = ResidualQuantizer(nq, d, "ema", K, algo=algo)
with torch.no_grad():
    m.codebooks.mul_(scale)
m = m.cuda().eval()
x = torch.randn(N, d, device="cuda")
torch.cuda.synchronize()
print("prepared", flush=True)
m._prepared(); torch.cuda.synchronize()
print("prep ok", flush=True)
with torch.no_grad():
    xq, idx, c = m(x)
torch.cuda.synchronize()
print("ok", float(c), int(idx.max()))
