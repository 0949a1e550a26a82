"""Reproduce on ONE GPU the codebooks an S-rank sharded C3 run converges to: the statistics of S different 1M-frame
shards are accumulated (the kernel adds into the caller's buffer) before ONE EMA refresh, exactly what the all-reduce
over S ranks produces.  Then time the encode kernel in that state and look at the certificate's counters and at the
code norms the error bound is built from.   python scripts/c3_shards_probe.py [S ...]"""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200 import _lib
from audio_generation_b200.quantizer import _ptr, _stream, RVQ_FLAG_COUNTERS


class A:
    algo, kernel = "tensor", "auto"


dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["c3"])
nq, K, d, N = wl["nq"], wl["K"], wl["d"], wl["frames"]
lib = _lib.load()
for S in [int(a) for a in sys.argv[1:]] or [1, 8]:
    q = bench.build_quantizer(wl, dev, A)
    shards = [torch.randn(N, d, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + r)) for r in range(S)]
    xq = torch.empty(N, d, device=dev)
    idx = torch.empty(N, nq, dtype=torch.int64, device=dev)
    csq = torch.empty(nq, dtype=torch.float64, device=dev)
    ws = q._workspace(dev)
    flat, ssum, scnt, rep = q._stats_buffers(dev)

    def encode(x, stats, flags=0):
        op, nrm, meta = q._prepared()
        _lib.check(lib.rvq_encode(_ptr(x), N, N, 0, d, 1, d, nq, K, _ptr(q.codebooks), _ptr(op), _ptr(nrm), _ptr(meta), _ptr(xq),
                                  _ptr(idx), _ptr(csq), _ptr(ssum if stats else None), _ptr(scnt if stats else None), _ptr(ws),
                                  ws.numel(), flags, _stream()), "rvq_encode")

    for step in range(150):
        flat[: ssum.numel() + scnt.numel()].zero_()
        for x in shards:
            encode(x, True)
        q._update_codebooks(shards[0], N, N, 0, d, 1, nq, idx, flat, ssum, scnt, rep)
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        flat[: ssum.numel() + scnt.numel()].zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        encode(shards[0], True)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    encode(shards[0], True, RVQ_FLAG_COUNTERS)
    torch.cuda.synchronize()
    c = q.read_counters()
    jobs = max(c[5], 1)
    norms = q.codebooks.norm(dim=2)                       # [nq, K]
    cnt = q.ema_count
    live = cnt > 0.5 * cnt.mean(dim=1, keepdim=True)
    ratio = (norms.max(dim=1).values / (norms * live).sum(1).div(live.sum(1).clamp(min=1))).tolist()
    dead = (~live).sum(1).tolist()
    print(f"S={S}: encode kernel (with statistics) {ms:.3f} ms after 150 updates; per tile-stage: two-candidate rows "
          f"{c[6] / jobs:.2f}, exact-scan rows {c[4] / jobs:.3f}", flush=True)
    print(f"   max ||c|| / mean live ||c|| per stage: {[round(r, 2) for r in ratio]}", flush=True)
    print(f"   codes with a count below half the stage mean: {dead}", flush=True)
