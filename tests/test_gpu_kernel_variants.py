"""GPU tests (-m gpu) of the two fused encode kernels and their launch variants.

d <= 128 runs `rvq_encode_tr_kernel` (residual resident in tensor memory, norm term folded into the MMA, cluster
multicast of the codebook stream); other d, or RVQ_KERNEL=tc, runs the generic `rvq_encode_tc_kernel`.  Both use the
same exact fp32 scorer for every frame the fp16 filter cannot certify, so on the same input they must return
bit-identical code indices and outputs, whatever the cluster size - and whether the update waits for the exact
re-rank or speculates on the approximate argmin and verifies / repairs afterwards (RVQ_SPEC=1).  The environment
switches are read once per process, hence the subprocesses.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys, torch
sys.path.insert(0, {root!r})
from audio_generation_b200 import ResidualQuantizer
nq, K, d, N, strided, update = {nq}, {K}, {d}, {N}, {strided}, {update}
torch.manual_seed(3)
m = ResidualQuantizer(nq, d, "ema", K)
with torch.no_grad():
    for q in range(nq):
        m.codebooks[q].mul_(0.7 ** q)
    m.ema_sum.copy_(m.codebooks)
m = m.cuda()
g = torch.Generator(device="cuda").manual_seed(5)
if strided:
    L = 125
    x = torch.randn(N // L, d, L, device="cuda", generator=g).permute(0, 2, 1)   # (B, L, d) view of (B, d, L)
else:
    x = torch.randn(N, d, device="cuda", generator=g)
m.train(update)
with torch.no_grad():
    xq, idx, commit = m(x, None, update_codebook=update)
torch.cuda.synchronize()
torch.save(dict(idx=idx.cpu(), xq=xq.contiguous().cpu(), commit=float(commit), cb=m.codebooks.cpu(),
                cnt=m.ema_count.cpu()), {out!r})
"""


def run_variant(tmp_path, name, env, **kw):
    out = str(tmp_path / f"{name}.pt")
    code = WORKER.format(root=ROOT, out=out, **kw)
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


@pytest.mark.parametrize("d,K,N,strided", [(128, 1024, 20000, False), (64, 300, 5000, False), (128, 512, 8000, True)])
def test_tmem_resident_kernel_equals_generic_kernel(tmp_path, d, K, N, strided):
    kw = dict(nq=5, K=K, d=d, N=N, strided=strided, update=False)
    a = run_variant(tmp_path, "tr2", {"RVQ_CLUSTER": "2"}, **kw)
    for name, env in [("tr1", {"RVQ_CLUSTER": "1"}), ("tr4", {"RVQ_CLUSTER": "4"}), ("spec", {"RVQ_SPEC": "1"}),
                      ("tc", {"RVQ_KERNEL": "tc"}), ("tc2", {"RVQ_KERNEL": "tc", "RVQ_CLUSTER_TC": "2"})]:
        b = run_variant(tmp_path, name, env, **kw)
        assert torch.equal(a["idx"], b["idx"]), name
        assert torch.equal(a["xq"], b["xq"]), name
        assert abs(a["commit"] - b["commit"]) <= 1e-6 * abs(a["commit"]), name


def test_statistics_variants_agree(tmp_path):
    """EMA statistics: bulk reductions (tr kernel) vs vector atomics (generic kernel) give the same update."""
    kw = dict(nq=3, K=512, d=128, N=30000, strided=False, update=True)
    a = run_variant(tmp_path, "tr", {}, **kw)
    b = run_variant(tmp_path, "tc", {"RVQ_KERNEL": "tc"}, **kw)
    assert torch.equal(a["idx"], b["idx"])
    assert torch.allclose(a["cnt"], b["cnt"], rtol=0, atol=0)
    assert torch.allclose(a["cb"], b["cb"], rtol=1e-5, atol=1e-6)


def test_many_stages_and_ragged_tail():
    """16 stages, N not a multiple of the 128-frame tile, more tiles than one wave of CTA pairs."""
    sys.path.insert(0, ROOT)
    from oracle import rvq_oracle as O
    from audio_generation_b200 import ResidualQuantizer
    torch.manual_seed(1)
    nq, K, d, N = 16, 256, 128, 148 * 128 * 2 + 77
    m = ResidualQuantizer(nq, d, "ema", K)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.8 ** q)
    m = m.cuda().eval()
    x = torch.randn(N, d, device="cuda")
    with torch.no_grad():
        xq, idx, commit = m(x)
    torch.cuda.synchronize()
    cbs = [m.codebooks[q].detach().cpu() for q in range(nq)]
    sub = torch.cat([torch.arange(0, 3000), torch.arange(N - 3000, N)])
    adj = O.adjudicate_indices(x[sub].cpu(), cbs, idx[sub].cpu())
    assert adj["n_illegal"] == 0, adj
    deq = sum(cbs[q][idx[sub, q].cpu()] for q in range(nq))
    assert (xq[sub].cpu() - deq).abs().max() <= 1e-5 * float(x.abs().max())
