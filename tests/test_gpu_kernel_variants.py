"""GPU tests (-m gpu) of the fused encode kernels and their launch variants.

d = 64 / 128 / 256 runs `rvq_encode_fr_kernel` (thread = frame, residual resident in tensor memory, norm term folded
into the MMA, cluster multicast of the codebook stream); other d, or kernel="generic", runs `rvq_encode_tc_kernel`;
kernel="tmem" is round 1's kernel for d <= 128 (separate scan and update warps).  All of them use the same exact
fp32 scorer for every frame the fp16 filter cannot certify, so on the same input they must return bit-identical code
indices and outputs, whatever the cluster size (for the "tmem" kernel cluster = 2 means the two CTAs of a cluster drive
their tensor cores as ONE `tcgen05.mma.cta_group::2` instruction stream).  The variants are selected through `flags` bits of `rvq_encode`
(include/rvq_sm100a.h), not through the environment.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_variant(kernel, cluster, nq, K, d, N, strided, update):
    from audio_generation_b200 import ResidualQuantizer
    torch.manual_seed(3)
    m = ResidualQuantizer(nq, d, "ema", K, kernel=kernel, cluster=cluster)
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
    return dict(idx=idx.cpu(), xq=xq.contiguous().cpu(), commit=float(commit), cb=m.codebooks.cpu(),
                cnt=m.ema_count.cpu())


@pytest.mark.parametrize("d,K,N,strided", [(128, 1024, 20000, False), (64, 300, 5000, False), (128, 512, 8000, True),
                                           (256, 1024, 12000, False), (256, 512, 6000, True)])
def test_kernels_agree_bitwise(d, K, N, strided):
    kw = dict(nq=5, K=K, d=d, N=N, strided=strided, update=False)
    a = run_variant("frame", 2, **kw)
    variants = [("frame", 1), ("frame", 4), ("generic", 0), ("generic", 2)]
    if d <= 128:   # independent CTAs (default), the cta_group::2 pair, one codebook stream multicast to four CTAs
        variants += [("tmem", 1), ("tmem", 2), ("tmem", 4)]
    for kernel, cluster in variants:
        b = run_variant(kernel, cluster, **kw)
        name = f"{kernel}/{cluster}"
        assert torch.equal(a["idx"], b["idx"]), name
        assert torch.equal(a["xq"], b["xq"]), name
        assert abs(a["commit"] - b["commit"]) <= 1e-6 * abs(a["commit"]), name


@pytest.mark.parametrize("d", [128, 256])
def test_statistics_variants_agree(d):
    """EMA statistics: the frame-resident kernel and the generic kernel give the same update."""
    kw = dict(nq=3, K=512, d=d, N=30000, strided=False, update=True)
    a = run_variant("frame", 0, **kw)
    b = run_variant("generic", 0, **kw)
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
