"""HostEncoder throughput against chunk size / buffers / frames per call (C2 shape): python scripts/e2e_tune.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200.quantizer import HostEncoder


class A:
    algo, kernel = "tensor", "auto"


wl = bench.WORKLOADS["c2"]
dev = torch.device("cuda", 0)
q = bench.build_quantizer(wl, dev, A).eval()
d, nq = wl["d"], wl["nq"]
for n_e2e in (1 << 19, 1 << 20):
    xh = torch.randn(n_e2e, d).pin_memory()
    for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19):
        for nbuf in (2, 3, 4):
            if chunk * nbuf > 2 * n_e2e:
                continue
            he = HostEncoder(q, chunk_frames=chunk, n_buffers=nbuf, packed=True)
            ih = torch.empty((n_e2e, he.bytes_per_frame(nq)), dtype=torch.uint8).pin_memory()
            for _ in range(3):
                he.encode(xh, ih)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                he.encode(xh, ih)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 8
            print(f"frames/call={n_e2e} chunk={chunk} buffers={nbuf}: {n_e2e / dt / 1e6:.1f} M frames/s  "
                  f"{n_e2e * d * 4 / dt / 1e9:.1f} GB/s H2D", flush=True)
            del he, ih
    del xh
