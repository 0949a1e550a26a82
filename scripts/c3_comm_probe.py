"""torchrun --nproc-per-node N scripts/c3_comm_probe.py : where does the sharded C3 step lose time?  Encode-kernel
and step time with (a) no all-reduce, (b) the all-reduce on the statistics buffer itself, (c) on a copy of it."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


class A:
    algo, kernel = "tensor", "auto"


wl = dict(bench.WORKLOADS["c3"])
q = bench.build_quantizer(wl, dev, A)
x = torch.randn(wl["frames"], wl["d"], device=dev)
tiny = torch.zeros(1, device=dev)
orig_update = q._update_codebooks
lockstep = [False]


def update(*a, **k):
    if lockstep[0]:                 # keeps the ranks in lockstep (4-byte all-reduce) without touching the statistics
        dist.all_reduce(tiny)
    return orig_update(*a, **k)


q._update_codebooks = update
WARM = int(os.environ.get("PROBE_WARM", "60"))
for name, sync, copy, lock in [("no all-reduce", False, False, False), ("all-reduce in place", True, False, False),
                               ("no all-reduce", False, False, False), ("4-byte all-reduce only (lockstep)", False, False, True),
                               ("all-reduce on a copy", True, True, False), ("all-reduce in place", True, False, False)]:
    q.sync_stats, q.comm_copy = sync, copy
    lockstep[0] = lock
    for _ in range(WARM):
        with torch.no_grad():
            q(x, None, update_codebook=True)
    torch.cuda.synchronize()
    dist.barrier()
    q.kernel_events, q.comm_events, q.update_events = [], [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        with torch.no_grad():
            q(x, None, update_codebook=True)
    e1.record()
    torch.cuda.synchronize()
    k = statistics.mean(a.elapsed_time(b) for a, b in q.kernel_events)
    c = statistics.mean(a.elapsed_time(b) for a, b in q.comm_events) if q.comm_events else 0.0
    u = statistics.mean(a.elapsed_time(b) for a, b in q.update_events)
    t = torch.tensor([e0.elapsed_time(e1) / 20, k, c, u], device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    q.kernel_events = q.comm_events = q.update_events = None
    if rank == 0:
        print(f"{name:34s} world={world}: " + " | ".join(
            f"rank {i}: step {a[0]:.3f} kernel {a[1]:.3f} allreduce {a[2]:.3f} maintenance {a[3]:.3f} ms" for i, a in enumerate(allt)),
            flush=True)
dist.destroy_process_group()
