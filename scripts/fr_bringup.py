"""Bring-up check of the frame-resident kernel (rvq_encode_fr.cu): bitwise against the exact-scan kernel and the
generic kernel on a ladder of shapes, event counters, and a quick timing.  python scripts/fr_bringup.py [quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_generation_b200 import ResidualQuantizer  # noqa: E402


def make(nq, K, d, kernel, cluster=0, algo="tensor"):
    torch.manual_seed(0)
    m = ResidualQuantizer(nq, d, "ema", K, kernel=kernel, cluster=cluster, algo=algo)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.7 ** q)
        m.ema_sum.copy_(m.codebooks)
    return m.cuda().eval()


def run(m, x, update=False):
    m.train(update)
    with torch.no_grad():
        out = m(x, None, update_codebook=update)
    torch.cuda.synchronize()
    return out


shapes = [(1, 256, 128, 128), (2, 256, 128, 300), (3, 512, 64, 1000), (4, 1024, 128, 5000), (3, 1024, 256, 3000),
          (8, 1024, 128, 148 * 256 + 77), (12, 1024, 256, 148 * 128 * 2 + 5), (2, 300, 128, 1029)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = shapes[:4]
if len(sys.argv) > 1 and sys.argv[1] == "time":
    shapes = []
ok = True
for nq, K, d, N in shapes:
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, d, device="cuda", generator=g)
    ref = make(nq, K, d, "auto", algo="exact_scan")
    xq_e, idx_e, c_e = run(ref, x)
    for kernel, cluster in [("frame", 1), ("frame", 2), ("generic", 0)]:
        m = make(nq, K, d, kernel, cluster)
        m.counters = True
        t0 = time.time()
        xq, idx, c = run(m, x)
        dt = time.time() - t0
        same_i = bool(torch.equal(idx, idx_e))
        same_x = bool(torch.equal(xq, xq_e))
        cerr = abs(float(c) - float(c_e)) / max(abs(float(c_e)), 1e-30)
        cnt = m.read_counters()[:3]
        nbad = int((idx != idx_e).any(dim=-1).sum())
        print(f"nq={nq} K={K} d={d} N={N} {kernel}/{cluster}: idx_equal={same_i} xq_equal={same_x} commit_rel={cerr:.1e} "
              f"bad_frames={nbad} counters={cnt} first_call_s={dt:.3f}", flush=True)
        ok &= same_i and same_x and cerr < 1e-6
# statistics
for d in (() if not shapes else (128, 256)):
    nq, K, N = 3, 512, 30000
    x = torch.randn(N, d, device="cuda")
    a = make(nq, K, d, "frame")
    b = make(nq, K, d, "generic")
    run(a, x, True)
    run(b, x, True)
    e_cnt = float((a.ema_count - b.ema_count).abs().max())
    e_cb = float((a.codebooks - b.codebooks).abs().max())
    print(f"stats d={d}: max|ema_count diff|={e_cnt:.3e} max|codebook diff|={e_cb:.3e}", flush=True)
    ok &= e_cnt == 0.0 and e_cb < 1e-4
# timing
for nq, K, d, N in [(8, 1024, 128, 1 << 20), (12, 1024, 256, 1 << 20)]:
    x = torch.randn(N, d, device="cuda")
    for kernel, cluster in (("frame", 2), ("frame", 1), ("tmem" if d <= 128 else "generic", 0)):
        m = make(nq, K, d, kernel, cluster)
        for _ in range(5):
            run(m, x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            with torch.no_grad():
                m(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = N * nq * 2 * K * d / (ms * 1e-3) / 1e12
        if kernel == "frame":
            m.counters = True
            run(m, x)
            c = m.read_counters()
            m.counters = False
            j = max(c[3], 1)
            print(f"  phases per job (cycles, warp 0 of each group): jobs={c[3]} scan={c[4] / j:.0f} (of which waiting for the "
                  f"accumulator {c[5] / j:.0f}) classify+rerank={c[6] / j:.0f} apply={c[7] / j:.0f} tail={c[8] / j:.0f}; "
                  f"reranked frames={c[0]} dirty={c[1]} winner != approximate argmin={c[2]}", flush=True)
            if c[10]:
                t = c[10]
                print(f"  MMA warp: total {t / 148:.0f} cycles per CTA; share waiting for the operand (a_ready) {c[11] / t:.2f}, for a "
                      f"free accumulator {c[12] / t:.2f}, for codebook data {c[13] / t:.2f}, issuing {c[14] / t:.2f}", flush=True)
        print(f"time nq={nq} K={K} d={d} N={N} {kernel}/{cluster}: {ms:.3f} ms  {N / ms / 1e3:.1f} M frames/s  {tf:.0f} TFLOP/s "
              f"({tf / 1635.7:.3f} of measured peak)", flush=True)
print("BRINGUP", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
