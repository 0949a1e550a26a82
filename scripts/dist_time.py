"""torchrun ... scripts/dist_time.py : where does a C3 step spend its time at N ranks?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from audio_generation_b200 import ResidualQuantizer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS["c3"]
nq, K, d, N = wl["nq"], wl["K"], wl["d"], int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = ResidualQuantizer(nq, d, "ema", K)
with torch.no_grad():
    m.codebooks.copy_(bench.synth_codebooks(nq, K, d))
    m.ema_sum.copy_(m.codebooks)
m = m.to(dev).train()
g = torch.Generator(device=dev).manual_seed(1234 + rank)
x = torch.randn(N, d, device=dev, generator=g)


def timed(fn, reps=5, warm_=2):
    for _ in range(warm_):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    t_full = timed(lambda: m(x, None, update_codebook=True), reps=10, warm_=warm)
    m.eval()
    t_enc = timed(lambda: m(x, None))
    m.train()
    flat = m._stats_buffers(dev)[0]
    t_ar = timed(lambda: dist.all_reduce(flat)) if world > 1 else 0.0
    t_zero = timed(lambda: flat.zero_())
    t_prep = timed(lambda: (m.invalidate(), m._prepared()))
t_all = torch.tensor([t_full, t_enc], device=dev)
if world > 1:
    gat = [torch.zeros_like(t_all) for _ in range(world)]
    dist.all_gather(gat, t_all)
    if rank == 0:
        print("per-rank [step(update), encode-only] ms:", [[round(float(v), 2) for v in g_] for g_ in gat])
if rank == 0:
    print(f"world={world} N/gpu={N}: step(update)={t_full:.3f} ms  encode-only={t_enc:.3f} ms  all_reduce({flat.numel()*4/1e6:.1f} MB)={t_ar:.3f} ms  "
          f"zero={t_zero:.3f} ms  prepare={t_prep:.3f} ms")
if world > 1:
    dist.destroy_process_group()
