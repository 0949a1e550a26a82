"""world_size-2 gloo test (CPU) of the only cross-shard step of the path: frames are sharded across ranks,
codebooks replicated, and ONE sum all-reduce of the EMA statistics makes every replica apply the same update
(SURVEY 8e).  The arithmetic here is the oracle's (no GPU in this container); the GPU path issues the same
torch.distributed.all_reduce on its flat [sum | cnt] buffer (audio_generation_b200/quantizer.py:_encode)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rvq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, cls="ema"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    nq, K, d, N = 3, 32, 16, 400
    m = O.ResidualQuantizerRef(nq, d, cls, K).train()
    x = torch.randn(N, d)
    shard = x[rank * N // world:(rank + 1) * N // world]
    with torch.no_grad():
        _, idx, _ = m(shard, None, update_codebook=True)       # all_reduce inside (dist is initialised)
    # flat statistics buffer layout used by the GPU path: [nq*K*d sums | nq*K counts], one all-reduce
    flat = torch.zeros(nq * K * d + nq * K)
    flat[rank::world] = 1.0
    dist.all_reduce(flat)
    # numpy arrays travel by value (a tensor travels as a file descriptor the exiting worker may already have closed)
    q.put((rank, m.codebooks.detach().numpy().copy(), m.ema_count.numpy().copy(), idx.numpy().copy(), float(flat.sum())))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("cls", ["ema", "base"])
def test_sharded_update_equals_single_process(cls):
    """"base" (gradient-trained codebooks): only the usage counts and the replacement vectors of stale codes cross the
    shards, in the same single all-reduce."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, cls)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    res = [(r, torch.from_numpy(cb), torch.from_numpy(cnt), torch.from_numpy(ix), tot) for r, cb, cnt, ix, tot in res]
    for p in procs:
        p.join(60)
    torch.manual_seed(7)
    nq, K, d, N = 3, 32, 16, 400
    ref = O.ResidualQuantizerRef(nq, d, cls, K).train()
    x = torch.randn(N, d)
    with torch.no_grad():
        _, ridx, _ = ref(x, None, update_codebook=True)
    assert torch.equal(res[0][1], res[1][1])                           # replicas stay bit-identical
    assert torch.allclose(res[0][1], ref.codebooks.detach(), rtol=1e-5, atol=1e-6)   # and equal the unsharded update
    assert torch.allclose(res[0][2], ref.ema_count, rtol=1e-6)
    assert torch.equal(torch.cat([res[0][3], res[1][3]]), ridx)        # encode needs no communication
    assert res[0][4] == nq * K * d + nq * K
