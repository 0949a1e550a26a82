"""torchrun --nproc-per-node N scripts/dist_check.py : sharded encode + EMA update over N GPUs equals the
unsharded update on one GPU (replicas bit-identical to each other; vs. single GPU within fp32-atomic tolerance)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from audio_generation_b200 import ResidualQuantizer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(11)
nq, K, d, N = 4, 1024, 256, 1 << 16


def force_stale(model):
    """A few far-away codes with a low EMA count: never selected, so the update must re-seed them - from frames that
    live on different ranks (the replacement vectors travel in the same all-reduce as the statistics)."""
    with torch.no_grad():
        # a dead 3 x 3 block of the 32 x 32 map: the SOM neighbourhood rescues its corners (two live neighbours each),
        # the centre and the edges stay below the cutoff
        for q in (0, 3):
            for y in (10, 11, 12):
                for xg in (10, 11, 12):
                    k = y * 32 + xg
                    model.codebooks[q, k] = 100.0
                    model.ema_sum[q, k] = 100.0
                    model.ema_count[q, k] = 0.05
    return model


m = force_stale(ResidualQuantizer(nq, d, "ema", K)).to(dev).train()        # defaults: SOM spreading + re-seeding on
x = torch.randn(N, d, device=dev)                      # same seed on every rank -> same full batch
shard = x[rank * N // world:(rank + 1) * N // world]
replaced = []
for step in range(2):
    with torch.no_grad():
        _, idx, commit = m(shard, None, update_codebook=True)
    replaced.append(m.n_replaced.tolist())
torch.cuda.synchronize()
cb = m.codebooks.clone()
gathered = [torch.empty_like(cb) for _ in range(world)]
dist.all_gather(gathered, cb)
identical = all(torch.equal(gathered[0], g) for g in gathered)
dist.barrier()
dist.destroy_process_group()                           # single-GPU reference below must not all-reduce
if rank == 0:
    torch.manual_seed(11)
    ref = force_stale(ResidualQuantizer(nq, d, "ema", K)).to(dev).train()
    ref_replaced = []
    for step in range(2):
        with torch.no_grad():
            ref(x, None, update_codebook=True)
        ref_replaced.append(ref.n_replaced.tolist())
    err = (ref.codebooks - cb).abs().max().item()
    cnt_err = (ref.ema_count - m.ema_count).abs().max().item()
    print(f"world={world} replicas_bit_identical={identical} max|codebook diff vs single GPU|={err:.3e} "
          f"max|ema_count diff|={cnt_err:.3e}")
    print(f"codes re-seeded per step and stage: sharded {replaced} single GPU {ref_replaced}")
    assert identical and err < 1e-3 and cnt_err < 1e-3
    assert replaced == ref_replaced and sum(replaced[0]) >= 4
