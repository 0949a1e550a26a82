"""torchrun --nproc-per-node N scripts/allreduce_probe.py : time of the statistics all-reduce alone (C3 payload)."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 12 * 1024 * 257
t = torch.ones(n, device=dev)
for _ in range(5):
    dist.all_reduce(t)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(t)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    ms = e0.elapsed_time(e1) / 20
    print(f"world={world} all_reduce of {n * 4 / 1e6:.1f} MB: {ms:.3f} ms per call "
          f"({2 * (world - 1) / world * n * 4 / ms / 1e6:.1f} GB/s bus), peer access 0->1: "
          f"{torch.cuda.can_device_access_peer(0, 1) if world > 1 else None}")
dist.destroy_process_group()
