"""GPU parity hardening (-m gpu): the committed golden fixtures through the CUDA path, the BASELINE configs at full
depth, inputs built to break the fp16 filter's certificate, the proven error bound measured, autograd with the
codebook maintenance in the same call, optimizer writes, and the sharded update under torchrun.

Everything that compares two GPU paths (tensor-core filter vs exact scan) is BITWISE: both use one exact fp32 scorer
(csrc/exact.cuh), so the certified filter may never change a winner.  Comparisons against the CPU oracle allow only
fp32 near-ties (oracle/rvq_oracle.py:adjudicate_indices).
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import rvq_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def quantizer(nq, K, d, cls="ema", **kw):
    from audio_generation_b200 import ResidualQuantizer
    return ResidualQuantizer(nq, d, cls, K, **kw)


def load_codebooks(m, cbs):
    with torch.no_grad():
        m.codebooks.copy_(torch.as_tensor(cbs))
        m.ema_sum.copy_(m.codebooks)
    return m


# --------------------------------------------------------------------------------------------- (i) golden fixtures
def test_golden_c1_reference_latents_through_cuda():
    """BASELINE configs[0]: the latents the UNMODIFIED reference encoder produced on networks/om.wav
    (tests/golden/make_golden.py), in the reference's own layout - a (1, 136, 512) view with strides (69632, 1, 136) of
    a (1, 512, 136) tensor (vae.py:313) - through the product path."""
    z = np.load(os.path.join(GOLD, "c1_reference.npz"))
    nq, K, d = [int(v) for v in z["ctor"]]
    torch.manual_seed(int(z["codebook_seed"]))
    cbs = torch.randn(nq, K, d) * float(z["codebook_scale"]) * torch.tensor([0.8 ** i for i in range(nq)])[:, None, None]
    if abs(float(cbs.double().abs().sum()) - float(z["codebook_checksum"])) > 1e-6 * float(z["codebook_checksum"]):
        pytest.skip("torch CPU RNG stream differs from the one that generated the fixture")
    m = load_codebooks(quantizer(nq, K, d, "base", vq_cutoff_freq=0.1, use_som=True), cbs).cuda().eval()
    x = torch.from_numpy(z["x_fp16"]).float()                               # (1, 136, 512)
    xs = x.permute(0, 2, 1).contiguous().cuda().permute(0, 2, 1)            # channel-major storage, frame view
    assert tuple(xs.shape) == tuple(z["x_shape"]) and not xs.is_contiguous()
    assert tuple(xs.stride()[1:]) == tuple(int(v) for v in z["x_stride"][1:])   # (.., 1, 136); dim 0 has size 1
    with torch.no_grad():
        xq, idx, commit = m(xs, None)
    assert idx.shape == tuple(z["ref_index_shape"]) and idx.dtype == torch.int64
    gold = torch.from_numpy(z["idx"].astype(np.int64))
    got = idx.reshape(-1, nq).cpu()
    adj = O.adjudicate_indices(x.reshape(-1, d), list(cbs), got)
    assert adj["n_illegal"] == 0, adj
    assert (got != gold).any(dim=1).sum() <= 1                              # fp32 near-ties only
    if torch.equal(got, gold):
        # "base": commit = commitment + codebook loss = 2 x the per-stage means the fixture holds
        assert abs(float(commit) - 2.0 * float(z["commit"].sum())) <= 1e-5 * 2.0 * float(z["commit"].sum())
        assert abs(float(xq.double().sum()) - float(z["xq_checksum"])) < 1e-3 * max(1.0, abs(float(z["xq_checksum"])))
    assert xq.stride()[1:] == xs.stride()[1:]                               # decoder-side rearrange stays a view


def test_golden_small_problem_through_cuda():
    """tests/golden/rvq_small.npz: indices, xq, residual, commit, EMA statistics and refreshed codebooks."""
    z = np.load(os.path.join(GOLD, "rvq_small.npz"))
    nq, K, d = z["codebooks"].shape
    m = load_codebooks(quantizer(nq, K, d, "ema", vq_cutoff_freq=0.0, use_som=False), z["codebooks"]).cuda().train()
    x = torch.from_numpy(z["x"]).cuda()
    with torch.no_grad():
        xq, idx, commit = m(x, None, update_codebook=True)
    assert torch.equal(idx.cpu(), torch.from_numpy(z["idx"]))
    assert np.allclose(xq.cpu().numpy(), z["xq"], atol=1e-5 * float(np.abs(z["x"]).max()))
    assert np.allclose((x - xq).cpu().numpy(), z["resid"], atol=2e-5 * float(np.abs(z["x"]).max()))
    assert abs(float(commit) - float(z["commit"].sum())) <= 1e-5 * float(z["commit"].sum())
    flat = m.last_stats.cpu().numpy()
    assert np.array_equal(flat[nq * K * d: nq * K * (d + 1)].reshape(nq, K), z["cnt"])          # counts exact
    assert np.allclose(flat[: nq * K * d].reshape(nq, K, d), z["sum"], rtol=1e-5, atol=1e-5)
    assert np.allclose(m.codebooks.cpu().numpy(), z["new_codebooks"], rtol=1e-4, atol=1e-5)


# --------------------------------------------------------------------------------------------- (ii) full depth
@pytest.mark.parametrize("name,nq,K,d,N", [("c3", 12, 1024, 256, 8192 + 37), ("c4", 32, 4096, 512, 8192),
                                           ("c2", 8, 1024, 128, 16384 + 5), ("default", 8, 1024, 512, 4000)])
def test_full_depth_configs_against_oracle(name, nq, K, d, N):
    """BASELINE configs[1..3] and the model default at their FULL stage depth and codebook size (frames reduced to what
    the CPU oracle finishes in seconds), teacher-forced; plus the exact-scan kernel bitwise."""
    torch.manual_seed(17)
    m = quantizer(nq, K, d)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.8 ** q)
    m = m.cuda().eval()
    x = torch.randn(N, d, device="cuda")
    with torch.no_grad():
        xq, idx, commit = m(x)
        m.algo = "exact_scan"
        xq_e, idx_e, commit_e = m(x)
    assert torch.equal(idx, idx_e) and torch.equal(xq, xq_e)
    cbs = [m.codebooks[q].detach().cpu() for q in range(nq)]
    adj = O.adjudicate_indices(x.cpu(), cbs, idx.cpu())
    assert adj["n_illegal"] == 0, adj
    assert adj["n_mismatch"] <= max(2, 2e-4 * idx.numel()), adj
    deq = sum(cbs[q][idx[:, q].cpu()] for q in range(nq))
    assert (xq.cpu() - deq).abs().max() <= 1e-5 * float(x.abs().max())


# --------------------------------------------------------------------------------------------- (iii) adversarial
def _adversarial_cases():
    g = torch.Generator().manual_seed(99)

    def randn(*s):
        return torch.randn(*s, generator=g)

    cases = []
    # codes that differ by ONE fp16 ulp of one coordinate (and pairs that differ by one fp32 ulp): every frame near
    # such a pair has two candidates whose fp16 operands are identical
    K, d = 512, 128
    cb = randn(K, d)
    cb[1::2] = cb[0::2]
    cb[1::2, 7] = (cb[0::2, 7].half().view(torch.int16) + 1).view(torch.half).float()
    cb[3::8] = cb[2::8]
    cb[3::8, 11] = torch.nextafter(cb[2::8, 11], torch.tensor(10.0))
    cases.append(("one-ulp pairs", torch.stack([cb, randn(K, d) * 0.5]), randn(4096, d)))
    # exact duplicates: lowest index must win; also a stage whose codes are ALL equal
    cb = randn(K, d)
    cb[K // 2:] = cb[: K // 2]
    cases.append(("duplicates", torch.stack([cb, torch.ones(K, d) * 0.25]), randn(3000, d)))
    # rank-1 codebook: every code a multiple of one direction (scores differ only through the norm term)
    u = randn(1, d)
    cb = u * torch.linspace(-2, 2, K)[:, None]
    cases.append(("rank-1", torch.stack([cb, cb * 0.01 + randn(K, d) * 1e-3]), randn(3000, d)))
    # tight clusters: 16 centres, members 1e-4 apart
    cen = randn(16, d)
    cb = cen.repeat_interleave(K // 16, 0) + randn(K, d) * 1e-4
    cases.append(("clusters", torch.stack([cb, randn(K, d) * 1e-4]), cen[torch.randint(0, 16, (3000,), generator=g)]
                  + randn(3000, d) * 1e-3))
    # heavy tails: Cauchy frames (huge dynamic range between frames and inside a frame)
    cauchy = torch.tan(math.pi * (torch.rand(3000, d, generator=g) - 0.5)).clamp(-1e6, 1e6)
    cases.append(("cauchy frames", torch.stack([randn(K, d), randn(K, d) * 0.3]), cauchy))
    # residuals far below the code scale: x ~ 2^-20, codes ~ 1 (late stages of a deep chain look like this)
    cases.append(("tiny residuals", torch.stack([randn(K, d), randn(K, d)]), randn(3000, d) * 2.0 ** -20))
    # ... and far above it
    cases.append(("huge frames", torch.stack([randn(K, d) * 2.0 ** -12, randn(K, d)]), randn(2000, d) * 2.0 ** 10))
    # K not a multiple of 16 / 128 / 256
    for Kodd in (37, 1000):
        cases.append((f"K={Kodd}", torch.stack([randn(Kodd, d), randn(Kodd, d) * 0.6]), randn(2500, d)))
    # zeros and constant frames
    xz = randn(1024, d)
    xz[::3] = 0.0
    xz[1::3] = 1.0
    cases.append(("zero / constant frames", torch.stack([randn(K, d), randn(K, d) * 0.5]), xz))
    return cases


@pytest.mark.parametrize("d,kernel", [(128, "auto"), (128, "frame"), (256, "auto"), (256, "frame"), (512, "auto"), (64, "auto")])
def test_certificate_survives_adversarial_inputs(d, kernel):
    """The fp16 filter + certificate must return exactly what the exact fp32 scan returns on inputs built to defeat
    it: identical fp16 operands, exact ties, degenerate geometry, heavy tails, extreme scale ratios, ragged K."""
    for name, cbs, x in _adversarial_cases():
        if d != 128:       # same constructions at another feature dimension: tile / truncate the features
            rep = (d + 127) // 128
            cbs = cbs.repeat(1, 1, rep)[:, :, :d].contiguous()
            x = x.repeat(1, rep)[:, :d].contiguous()
        nq, K, _ = cbs.shape
        m = load_codebooks(quantizer(nq, K, d, kernel=kernel), cbs).cuda().eval()
        xc = x.cuda()
        with torch.no_grad():
            m.algo = "tensor"
            xq_t, idx_t, c_t = m(xc)
            m.algo = "exact_scan"
            xq_e, idx_e, c_e = m(xc)
        assert torch.equal(idx_t, idx_e), (name, d, kernel, int((idx_t != idx_e).sum()))
        assert torch.equal(xq_t, xq_e), (name, d, kernel)
        # and the exact scan itself is the fp32 argmin up to summation order (duplicates: lowest index wins)
        adj = O.adjudicate_indices(x, [cbs[q] for q in range(nq)], idx_e.cpu())
        assert adj["n_illegal"] == 0, (name, adj)
        if name == "duplicates":
            assert int(idx_e[:, 0].max()) < K // 2 and int(idx_e[:, 1].max()) == 0


# --------------------------------------------------------------------------------------------- (vi) the error bound
@pytest.mark.parametrize("K,d", [(1024, 128), (1024, 256), (512, 512), (512, 64)])
def test_proven_error_bound_holds_on_every_score(K, d):
    """|approximate - exact| <= E_k for EVERY score the filter produces (~10^7 per shape, several
    input distributions), E_k being the bound with the code's OWN norm
        E_k = 1.02 * 2^-9 * rs * cs_k + 2^-15 * (na * cs_k^2 + 2 rs cs_k) + d * 2^-14        (scaled units)
    rs = 2^a ||r||, cs_k = 2^b ||c_k||, na = 2^(a-b).  The frame's threshold is built from the stage's norm CAP cs0
    (cb_meta[1]); codes above it have their allowance X_k = rs xc_k + na X2_k subtracted by the filter (optimistic scores),
    so the hook returns v = s~ - X_k and the test checks |v + X_k - exact| <= E_k.  Measured through the bring-up hook of
    the generic kernel; the "outliers" codebook makes sure codes above the cap exist."""
    from audio_generation_b200 import _lib
    from audio_generation_b200.quantizer import _ptr, _stream
    lib = _lib.load()
    Kpad = (K + 255) // 256 * 256
    g = torch.Generator(device="cuda").manual_seed(5)
    worst = 0.0
    n_scores = n_large = 0
    dists = ["gauss", "gauss_small", "uniform", "sparse", "cauchy", "aligned", "outliers"]
    reps = max(1, int(1e8 / (len(dists) * 128 * K)) // 8)
    gamma = (d // 8 + 4) * 2.0 ** -24
    c1 = (1.02 * 2.0 ** -9 + 2 * 2.0 ** -15 + 2 * gamma) * 1.01
    c2 = (2.0 ** -15 + gamma) * 1.01
    for dist in dists:
        m = quantizer(1, K, d)
        with torch.no_grad():
            if dist == "uniform":
                m.codebooks.uniform_(-1, 1)
            elif dist == "sparse":
                m.codebooks.mul_((torch.rand_like(m.codebooks) < 0.1).float())
            elif dist == "outliers":          # a trained codebook: most codes small, a tenth kept their initial scale
                m.codebooks.mul_(0.2)
                m.codebooks[0, ::10] *= 20.0
        m = m.cuda()
        op, nrm, meta = m._prepared()
        sb = float(meta.reshape(1, 8)[0, 0])
        cs0 = float(meta.reshape(1, 8)[0, 1]) * sb                      # the stage's norm cap (scaled)
        cb64 = m.codebooks[0].double()
        csk = cb64.norm(dim=1) * sb * (1 + 1e-5)                        # [K]
        large = csk > cs0
        n_large += int(large.sum())
        dcs = torch.where(large, csk - cs0, torch.zeros_like(csk))
        dcs2 = torch.where(large, csk * csk - cs0 * cs0, torch.zeros_like(csk))
        xc_dev = nrm[Kpad * 9: Kpad * 10][:K].double()                  # NormLayout: xc behind norms + slices
        # (the device derives cs_k and the cap in fp32: codes barely above the cap differ in the last digits)
        assert torch.allclose(xc_dev, c1 * dcs, rtol=1e-2, atol=c1 * cs0 * 1e-5)
        for rep in range(reps):
            x = torch.randn(128, d, device="cuda", generator=g)
            if dist == "gauss_small":
                x = x * 1e-3
            elif dist == "cauchy":
                x = torch.tan(math.pi * (torch.rand(128, d, device="cuda", generator=g) - 0.5)).clamp(-1e4, 1e4)
            elif dist == "aligned":      # frames ON codes: scores near their minimum, worst case for cancellation
                x = m.codebooks[0][torch.randint(0, K, (128,), device="cuda", generator=g)] * (1 + 1e-3 * x)
            scores = torch.empty((128, Kpad), device="cuda")
            rs_out = torch.zeros(128, device="cuda")
            _lib.check(lib.rvq_debug_stage_scores(_ptr(x), d, K, 0, _ptr(op), _ptr(nrm), _ptr(meta), _ptr(scores),
                                                  _ptr(rs_out), _stream()), "rvq_debug_stage_scores")
            sa = rs_out.double()                                           # 2^a per frame
            exact = sa[:, None] * sb * ((cb64 * cb64).sum(1)[None, :] - 2.0 * x.double() @ cb64.t())
            rs = x.double().norm(dim=1) * 1.00002 * sa
            na = sa / sb
            Ek = (1.02 * 2.0 ** -9 * rs[:, None] * csk[None, :] +
                  2.0 ** -15 * (na[:, None] * (csk * csk)[None, :] + 2.0 * rs[:, None] * csk[None, :]) + d * 2.0 ** -14)
            Xk = rs[:, None] * xc_dev[None, :] + na[:, None] * (c2 * dcs2)[None, :]
            ratio = ((scores[:, :K].double() + Xk - exact).abs() / Ek).max()
            worst = max(worst, float(ratio))
            # and the allowance really covers what the own-norm bound exceeds the cap-based bound by
            E0 = (1.02 * 2.0 ** -9 * rs * cs0 + 2.0 ** -15 * (na * cs0 * cs0 + 2.0 * rs * cs0) + d * 2.0 ** -14)
            E32k = (d // 8 + 4) * 2.0 ** -24 * (na[:, None] * (csk * csk)[None, :] + 2.0 * rs[:, None] * csk[None, :])
            E32_0 = (d // 8 + 4) * 2.0 ** -24 * (na * cs0 * cs0 + 2.0 * rs * cs0)
            need = (Ek + E32k) - (E0 + E32_0)[:, None]
            assert bool((Xk[:, large] >= need[:, large] - 1e-6 * E0[:, None]).all())
            n_scores += 128 * K
    assert worst < 1.0, worst
    assert n_scores >= 1e7 and n_large > 0


# --------------------------------------------------------------------------------------------- autograd + maintenance
@pytest.mark.parametrize("cls", ["ema", "base"])
def test_gradients_match_oracle_with_codebook_update_in_the_same_call(cls):
    """update_codebook=True rewrites the codebooks (EMA refresh, re-seeding of stale codes - one of them deliberately a
    code that WAS hit) before backward runs: the gradient must still belong to the codebooks the indices and the
    returned commit loss were computed with."""
    torch.manual_seed(2)
    nq, K, d, B, L = 3, 64, 64, 4, 100
    kw = dict(vq_cutoff_freq=0.9, use_som=True, som_kernel_type="hard")
    m = quantizer(nq, K, d, cls, **kw)
    ref = O.ResidualQuantizerRef(nq, d, cls, K, **kw)
    with torch.no_grad():
        m.ema_count.fill_(0.3)             # below the cutoff even after one hit-weighted EMA step: codes get re-seeded
        ref.load_state_dict({k: v.clone() for k, v in m.state_dict().items() if k in ref.state_dict()}, strict=False)
    m = m.cuda().train()
    ref.train()
    xc = torch.randn(B, d, L)
    x_ref = xc.permute(0, 2, 1).clone().requires_grad_(True)
    x_gpu = xc.cuda().permute(0, 2, 1).requires_grad_(True)              # the reference's strided view
    w = torch.randn(B, L, d)
    xq_r, idx_r, c_r = ref(x_ref, None, update_codebook=True)
    ((xq_r * w).sum() + 3.0 * c_r).backward()
    xq_g, idx_g, c_g = m(x_gpu, None, update_codebook=True)
    ((xq_g * w.cuda()).sum() + 3.0 * c_g).backward()
    assert torch.equal(idx_g.cpu(), idx_r)
    assert int(m.n_replaced.sum()) > 0 and m.n_replaced.tolist() == ref.n_replaced
    assert abs(float(c_g) - float(c_r)) <= 1e-5 * abs(float(c_r))
    gx = x_gpu.grad.cpu()
    assert torch.allclose(gx, x_ref.grad, rtol=1e-4, atol=1e-6 * float(x_ref.grad.abs().max()))
    if cls == "base":
        gc = m.codebooks.grad.cpu()
        assert torch.allclose(gc, ref.codebooks.grad, rtol=1e-4, atol=1e-6 * float(ref.codebooks.grad.abs().max()))
        # usage counts are kept for "base" too, so the stale-cluster report means something (training.py:435,461)
        assert torch.allclose(m.ema_count.cpu(), ref.ema_count, rtol=1e-6)
    assert torch.allclose(m.codebooks.detach().cpu(), ref.codebooks.detach(), rtol=1e-4, atol=1e-5)
    assert m.get_stale_clusters() == ref.get_stale_clusters()


@pytest.mark.parametrize("opt_kw", [dict(foreach=True), dict(fused=True), dict(foreach=False)])
def test_optimizer_writes_are_seen_without_invalidate(opt_kw):
    """quantizer_class="base": the optimizer rewrites the codebooks in place; the derived fp16 operands must follow
    WITHOUT anyone calling invalidate() (training.py:385-390 never will)."""
    torch.manual_seed(4)
    nq, K, d = 3, 256, 128
    m = quantizer(nq, K, d, "base", vq_cutoff_freq=0.0, use_som=False).cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-2, **opt_kw)
    x = torch.randn(2000, d, device="cuda")
    for step in range(3):
        opt.zero_grad()
        xq, idx, commit = m(x, None, update_codebook=True)
        commit.backward()
        opt.step()
        # a fresh module holding the stepped weights is the truth for the NEXT call
        fresh = quantizer(nq, K, d, "base", vq_cutoff_freq=0.0, use_som=False, algo="exact_scan").cuda().eval()
        with torch.no_grad():
            fresh.codebooks.copy_(m.codebooks)
            _, idx_f, _ = fresh(x)
            m.eval()
            _, idx_m, _ = m(x)
            m.train()
        assert torch.equal(idx_m, idx_f), (opt_kw, step)


def test_dequantize_into_reference_layout_with_ragged_tail():
    """rvq_dequantize with a (B, d, L)-backed output and N % 128 != 0 (the strided branch dropped
    features of the last group of 128 frames)."""
    import ctypes as C
    from audio_generation_b200 import _lib
    from audio_generation_b200.quantizer import _ptr, _stream
    lib = _lib.load()
    nq, K, d, B, L = 3, 64, 64, 2, 50                     # N = 100
    m = quantizer(nq, K, d).cuda().eval()
    idx = torch.randint(0, K, (B * L, nq), device="cuda")
    out = torch.full((B, d, L), float("nan"), device="cuda")
    _lib.check(lib.rvq_dequantize(_ptr(m.codebooks), _ptr(idx), B * L, L, d * L, 1, L, d, 0, nq, K, None, 0, _ptr(out),
                                  _stream()), "rvq_dequantize")
    want = sum(m.codebooks[q][idx[:, q]] for q in range(nq)).reshape(B, L, d).permute(0, 2, 1)
    assert not torch.isnan(out).any()
    assert torch.allclose(out, want, atol=1e-6)


def test_host_encoder_result_is_complete_on_return_and_packed_mode():
    """HostEncoder.encode: the host may read the codes as soon as the call returns (no extra synchronize), and
    packed=True returns the wire format."""
    from audio_generation_b200.quantizer import HostEncoder
    m = quantizer(8, 1024, 128).cuda().eval()
    xh = torch.randn(300000, 128).pin_memory()
    with torch.no_grad():
        _, idx, _ = m(xh.cuda())
    want = idx.cpu()
    for _ in range(3):
        ih = HostEncoder(m, chunk_frames=1 << 15).encode(xh)
        assert torch.equal(ih, want)                       # read immediately: no torch.cuda.synchronize() here
    he = HostEncoder(m, chunk_frames=1 << 15, packed=True)
    ph = he.encode(xh)
    assert ph.dtype == torch.uint8 and ph.shape == (300000, 10) and he.bytes_per_frame() == 10
    assert torch.equal(m.unpack_indices(ph.cuda()).cpu(), want)


def test_unaligned_contiguous_input_is_copied_not_refused():
    """A contiguous view at an odd offset, and a single-frame (B, d, 1) -> (B, 1, d) permute (contiguous, stride(1)=1):
    both must be encoded."""
    m = quantizer(2, 64, 64).cuda().eval()
    big = torch.randn(64 * 101 + 1, device="cuda")
    x = big[1:].view(101, 64)                              # 4-byte aligned only
    xs = torch.randn(5, 64, 1, device="cuda").permute(0, 2, 1)
    with torch.no_grad():
        _, i1, _ = m(x)
        _, i2, _ = m(x.clone())
        _, i3, _ = m(xs)
        _, i4, _ = m(xs.contiguous().clone())
    assert torch.equal(i1, i2) and torch.equal(i3, i4)


# --------------------------------------------------------------------------------------------- (v) torchrun
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_update_two_ranks_torchrun():
    """What scripts/dist_check.py asserts, under pytest: after sharded updates with the SOM neighbourhood and stale-code
    re-seeding on, replicas are bit-identical and match the single-GPU update."""
    env = dict(os.environ)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "scripts", "dist_check.py")], capture_output=True, text=True, timeout=600,
                       env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "replicas_bit_identical=True" in r.stdout


@pytest.mark.parametrize("d,L", [(512, 150), (128, 500)])
def test_encode_call_is_cuda_graph_capturable(d, L):
    """The eval-mode call (encode only) on the reference's strided latent view, captured once in a CUDA graph and
    replayed on new inputs: same codes as the eager call (nothing on the path synchronises or allocates outside the
    stream-ordered allocator)."""
    m = quantizer(6, 512, d).cuda().eval()
    static_x = torch.randn(4, d, L, device="cuda")
    xv = static_x.permute(0, 2, 1)
    with torch.no_grad():
        m(xv)                                             # builds the operands and the workspace outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            xq_g, idx_g, commit_g = m(xv)
        for seed in (1, 2):
            torch.manual_seed(seed)
            static_x.copy_(torch.randn(4, d, L, device="cuda"))
            g.replay()
            torch.cuda.synchronize()
            xq_e, idx_e, commit_e = m(xv)
            assert torch.equal(idx_g, idx_e) and torch.equal(xq_g, xq_e)
            assert abs(float(commit_g) - float(commit_e)) <= 1e-6 * abs(float(commit_e))


# ------------------------------------------------------------------- partly filled tiles (small calls) and (B, d, L) storage
@pytest.mark.parametrize("d", [256, 512])
@pytest.mark.parametrize("layout", ["rows", "bdl"])
def test_small_calls_with_partly_filled_tiles_are_bit_identical(d, layout):
    """Calls below 148 x 128 frames run tiles of ceil(N / SMs) frames rounded to 32 (rvq_encode_tc.cu: tile_rows): every
    tile size from one pass to the full tile, frame counts on both sides of each boundary, in the row layout and in the
    reference's (B, d, L) storage, against the exact-scan kernel (bitwise: it never uses tiles of this kind) - encode
    and the EMA statistics of an update step."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    nq, K = 4, 300
    torch.manual_seed(3)
    m = quantizer(nq, K, d, vq_cutoff_freq=0.0, use_som=False).cuda()
    ref = quantizer(nq, K, d, vq_cutoff_freq=0.0, use_som=False, algo="exact_scan").cuda()
    ref.load_state_dict(m.state_dict())
    for N in [1, 31, 33, sms * 32 - 1, sms * 32 + 1, sms * 64 + 5, sms * 96 + 1, sms * 128 - 1, sms * 128 + 7]:
        if layout == "rows":
            x = torch.randn(N, d, device="cuda")
        else:
            B = 3 if N % 3 == 0 and N >= 3 else 1
            x = torch.randn(B, d, N // B, device="cuda").permute(0, 2, 1)      # vae.py:313: a VIEW, features strided
            assert not x.is_contiguous() or N == 1
        m.eval(), ref.eval()
        with torch.no_grad():
            xq, idx, commit = m(x)
            xq_e, idx_e, commit_e = ref(x)
        assert torch.equal(idx, idx_e), (N, layout)
        assert torch.equal(xq, xq_e), (N, layout)
        assert xq.stride() == x.stride() or N == 1
    # one update step on a partly filled tiling: same statistics, same refreshed codebooks
    m.train(), ref.train()
    x = torch.randn(2, d, 150, device="cuda").permute(0, 2, 1) if layout == "bdl" else torch.randn(300, d, device="cuda")
    with torch.no_grad():
        _, idx, _ = m(x, None, update_codebook=True)
        _, idx_e, _ = ref(x, None, update_codebook=True)
    assert torch.equal(idx, idx_e)
    assert torch.equal(m.ema_count, ref.ema_count)
    assert torch.allclose(m.codebooks, ref.codebooks, rtol=1e-5, atol=1e-6)       # float atomics: summation order only


def test_prepared_allowance_tables_are_consistent():
    """K0 (rvq_aux.cu:k0_bound): the byte table bounds the fp32 allowance factor of every code from above
    (xb_k * U1 >= xc_k), codes at or below the stage's norm cap carry no allowance, the chunk flags mark exactly the
    chunks that hold one, padding codes none - on a codebook with outliers (cap = lower quartile) and on a freshly
    initialised one (norms within 25 %: cap = the maximum, no allowance at all)."""
    nq, K, d = 3, 700, 256
    Kpad = (K + 255) // 256 * 256
    for outliers in (True, False):
        torch.manual_seed(5)
        m = quantizer(nq, K, d)
        with torch.no_grad():
            m.codebooks.copy_(torch.randn(nq, K, d))
            if outliers:
                m.codebooks[:, ::17] *= 4.0
            m.ema_sum.copy_(m.codebooks)
        m = m.cuda().eval()
        op, nrm, meta = m._prepared()
        torch.cuda.synchronize()
        meta = meta.reshape(nq, 8).cpu()
        xc = nrm[9 * nq * Kpad:10 * nq * Kpad].reshape(nq, Kpad).cpu()
        raw = nrm[10 * nq * Kpad:].view(torch.uint8)
        xb = raw[:nq * Kpad].reshape(nq, Kpad).cpu().float()
        flag = raw[nq * Kpad:nq * Kpad + 4 * nq * (Kpad // 256)].view(torch.int32).reshape(nq, Kpad // 256).cpu()
        for q in range(nq):
            sb, cap, U2, cnmax, U1 = [float(meta[q, i]) for i in (0, 1, 5, 6, 7)]
            cs = m.codebooks[q].detach().cpu().double().norm(dim=1) * sb
            large = cs > cap * sb * (1 + 1e-4)
            small = cs < cap * sb * (1 - 1e-4)
            assert (xc[q, K:] == 0).all() and (xb[q, K:] == 0).all()
            assert (xb[q] * U1 >= xc[q]).all()                          # the table is an upper bound ...
            assert (xb[q] * U1 <= xc[q] + U1 * 2.001).all()             # ... within two units
            assert (xc[q, :K][small] == 0).all() and (xc[q, :K][large] > 0).all()
            assert torch.equal(flag[q] != 0, (xc[q].reshape(-1, 256) > 0).any(dim=1))
            if outliers:
                assert cap < cnmax * 0.5 and large.sum() >= K // 17     # capped at the live codes' scale
            else:
                assert cap == pytest.approx(cnmax, rel=1e-4) and not large.any() and U1 == 0
