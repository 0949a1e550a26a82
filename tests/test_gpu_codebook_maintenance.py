"""GPU parity of the codebook maintenance kernels through the C ABI: SOM neighbourhood
spreading of the EMA statistics and stale-code re-seeding, against oracle/rvq_oracle.py.  Bars: BIT-EXACT for the
stencil (same term order, separate fp32 multiply/add), for the replacement vectors (same fp32 subtraction chain as
the encode kernel) and for the re-seeded state; the module-level check inherits the EMA tolerances (atomics)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import rvq_oracle as O

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("kernel,t", [("hard", 0), ("hard", 7), ("gaussian", 0), ("gaussian", 30)])
def test_som_spread_bit_exact(kernel, t):
    from audio_generation_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(5)
    sizes = [512, 300, 64, 7]                       # ragged: maps 16x32, 15x20, 8x8, 1x7; K padded to 512
    nq, K, d = len(sizes), 512, 192
    sm = torch.randn(nq, K, d, device="cuda")
    cnt = torch.randint(0, 50, (nq, K), device="cuda").float()
    osm, ocnt = torch.full_like(sm, -7.0), torch.full_like(cnt, -7.0)
    radius, w = O.som_weights(kernel, t)
    hw = []
    for k in sizes:
        hw += list(O.approximate_square_root(k))
    wf = [float(v) for v in np.asarray(w).reshape(-1)]
    _lib.check(lib.rvq_som_spread(_p(sm), _p(cnt), _p(osm), _p(ocnt), (C.c_int * len(hw))(*hw), nq, K, d, radius,
                                  (C.c_float * len(wf))(*wf), _s()), "rvq_som_spread")
    torch.cuda.synchronize()
    for q, k in enumerate(sizes):
        h, wd = O.approximate_square_root(k)
        rsm, rcnt = O.som_spread_ref(sm[q].cpu().numpy(), cnt[q].cpu().numpy(), h, wd, radius, w)
        assert np.array_equal(osm[q].cpu().numpy(), rsm), (q, kernel)
        assert np.array_equal(ocnt[q].cpu().numpy(), rcnt), (q, kernel)
    # in-place and oversized windows are refused
    assert lib.rvq_som_spread(_p(sm), _p(cnt), _p(sm), _p(ocnt), (C.c_int * len(hw))(*hw), nq, K, d, radius,
                              (C.c_float * len(wf))(*wf), _s()) == -1
    assert lib.rvq_som_spread(_p(sm), _p(cnt), _p(osm), _p(ocnt), (C.c_int * len(hw))(*hw), nq, K, d, 5,
                              (C.c_float * len(wf))(*wf), _s()) == -1


@pytest.mark.parametrize("strided", [False, True])
def test_reseed_gather_and_apply_bit_exact(strided):
    from audio_generation_b200 import ResidualQuantizer, _lib
    lib = _lib.load()
    torch.manual_seed(9)
    nq, K, d, B, L = 4, 256, 128, 3, 200
    N = B * L
    m = ResidualQuantizer(nq, d, "ema", K, vq_cutoff_freq=0).cuda().eval()
    xs = torch.randn(B, d, L, device="cuda")
    x = xs.permute(0, 2, 1) if strided else xs.permute(0, 2, 1).contiguous()      # (B, L, d)
    with torch.no_grad():
        _, idx, _ = m(x)
    idx2 = idx.reshape(N, nq).contiguous()
    cbs = [m.codebooks[q].cpu() for q in range(nq)]
    res = O.stage_residuals_from_indices(x.reshape(N, d).cpu(), cbs, idx2.cpu())
    seed = 0xC0FFEE1234
    rep = torch.full((nq, K, d), 3.0, device="cuda")
    # two "ranks" holding the two halves of the frames: the sum of their replacement buffers is the single-rank one
    half = L if strided else N // 2                                 # strided: rank 0 = batch item 0 only
    parts = []
    for off, n_loc in ((0, half), (half, N - half)):
        xb = x.reshape(N, d)[off:off + n_loc] if not strided else x[off // L:(off + n_loc) // L]
        ib = idx2[off:off + n_loc].contiguous()
        r_ = torch.full((nq, K, d), 3.0, device="cuda")
        if strided:
            Lb, sb_, sl_, sd_ = L, xb.stride(0), xb.stride(1), xb.stride(2)
        else:
            xb = xb.contiguous()
            Lb, sb_, sl_, sd_ = n_loc, 0, d, 1
        _lib.check(lib.rvq_reseed_gather(_p(xb), n_loc, Lb, sb_, sl_, sd_, d, nq, K, _p(m.codebooks), _p(ib), None,
                                         0.99, 1.0, seed, off, N, _p(r_), _s()), "rvq_reseed_gather")
        parts.append(r_)
    torch.cuda.synchronize()
    total = parts[0] + parts[1]
    for q in range(nq):
        ref = O.reseed_vectors_ref(res[q], q, K, seed)
        assert torch.equal(total[q].cpu(), ref), q
        assert lib.rvq_reseed_frame(seed, q, K, 3, N) == O.reseed_frame_ref(seed, q, K, 3, N)
    # with ema_count given, only codes that can fall below the cutoff are gathered
    cnt = torch.full((nq, K), 5.0, device="cuda")
    cnt[1, 10] = 0.3
    cnt[2, 200] = 1.0                                                 # 0.99 * 1.0 < 1.0: may become stale
    x2 = x.reshape(N, d).contiguous()
    _lib.check(lib.rvq_reseed_gather(_p(x2), N, N, 0, d, 1, d, nq, K, _p(m.codebooks), _p(idx2), _p(cnt), 0.99, 1.0,
                                     seed, 0, N, _p(rep), _s()), "rvq_reseed_gather")
    torch.cuda.synchronize()
    nz = (rep != 0).any(dim=2).nonzero().tolist()
    assert nz == [[1, 10], [2, 200]]
    assert torch.equal(rep[1, 10], total[1, 10]) and torch.equal(rep[2, 200], total[2, 200])
    # apply: stale codes (count < cutoff) take the replacement; the rest is untouched
    cb = m.codebooks.detach().clone()
    es = torch.randn_like(cb)
    cnt[2, 200] = 0.999
    cb0, es0, cnt0 = cb.clone(), es.clone(), cnt.clone()
    nrep = torch.full((nq,), -1, dtype=torch.int32, device="cuda")
    _lib.check(lib.rvq_reseed_apply(_p(cb), _p(cnt), _p(es), _p(total), None, nq, K, d, 1.0, 0.75, _p(nrep), _s()),
               "rvq_reseed_apply")
    torch.cuda.synchronize()
    assert nrep.tolist() == [0, 1, 1, 0]
    for q in range(nq):
        rcb, rc, rs, _ = O.reseed_apply_ref(cb0[q].cpu(), cnt0[q].cpu(), es0[q].cpu(), total[q].cpu(), 1.0, 0.75)
        assert torch.equal(cb[q].cpu(), rcb) and torch.equal(cnt[q].cpu(), rc) and torch.equal(es[q].cpu(), rs)


@pytest.mark.parametrize("kernel", ["hard", "gaussian"])
def test_module_update_with_som_and_reseed_matches_oracle(kernel):
    """Whole update step of the drop-in module (encode + statistics -> SOM -> EMA -> re-seed) against the oracle
    module over several steps, with codes forced stale."""
    from audio_generation_b200 import ResidualQuantizer
    torch.manual_seed(11)
    nq, K, d, N = 3, 256, 128, 6000
    m = ResidualQuantizer(nq, d, "ema", K, vq_cutoff_freq=1.0, use_som=True, som_kernel_type=kernel, reseed_seed=42)
    ref = O.ResidualQuantizerRef(nq, d, "ema", K, vq_cutoff_freq=1.0, use_som=True, som_kernel_type=kernel,
                                 reseed_seed=42)
    with torch.no_grad():
        for q in range(nq):
            m.codebooks[q].mul_(0.7 ** q)
        m.codebooks[0, 17] = 50.0                        # never selected: goes stale and must be re-seeded
        m.codebooks[2, 255] = -50.0
        m.ema_sum.copy_(m.codebooks)
        m.ema_count[0, 17] = 0.2
        m.ema_count[2, 255] = 0.2
        ref.codebooks.copy_(m.codebooks)
        ref.ema_sum.copy_(m.ema_sum)
        ref.ema_count.copy_(m.ema_count)
    m = m.cuda().train()
    ref.train()
    assert m.get_stale_clusters() == [1, 0, 1] == ref.get_stale_clusters()
    for step in range(3):
        x = torch.randn(N, d, device="cuda")
        with torch.no_grad():
            _, idx, _ = m(x, None, update_codebook=True)
            _, ridx, _ = ref(x.cpu(), None, update_codebook=True)
        torch.cuda.synchronize()
        assert (idx.cpu() == ridx).all(dim=1).float().mean() > 0.999
        assert m.n_replaced.tolist() == ref.n_replaced
        if step == 0:       # (0, 17) is rescued by its four neighbours' hits (sigma_0 = 1); the corner code is not
            assert ref.n_replaced[2] == 1 and torch.equal(m.codebooks[2, 255].cpu(), ref.codebooks[2, 255])
        assert torch.allclose(m.ema_count.cpu(), ref.ema_count, rtol=1e-5, atol=1e-5)
        assert torch.allclose(m.ema_sum.cpu(), ref.ema_sum, rtol=1e-4, atol=1e-4)
        assert torch.allclose(m.codebooks.cpu(), ref.codebooks.detach(), rtol=1e-4, atol=1e-4)
    assert int(m.update_steps) == 3 == int(ref.update_steps)
    assert m.get_stale_clusters() == ref.get_stale_clusters()
    # the re-seeded codes are live data now: they were selected in later steps
    assert float(m.ema_count[0, 17]) >= 1.0
    # state survives a checkpoint round trip, including the step counter that drives sigma_t and the seed
    m2 = ResidualQuantizer(nq, d, "ema", K, vq_cutoff_freq=1.0, use_som=True, som_kernel_type=kernel, reseed_seed=42)
    m2.load_state_dict(m.state_dict())
    m2 = m2.cuda().train()
    x = torch.randn(N, d, device="cuda")
    with torch.no_grad():
        m(x, None, update_codebook=True)
        m2(x, None, update_codebook=True)
    assert int(m2.update_steps) == 4
    assert torch.allclose(m.codebooks, m2.codebooks, rtol=1e-4, atol=1e-5)


def test_kernels_reproduce_the_golden_fixture():
    """The committed vectors of tests/golden/codebook_maint.npz through the C ABI, bit for bit."""
    import os
    from audio_generation_b200 import _lib
    lib = _lib.load()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codebook_maint.npz"))
    K, d, N, q, seed = int(z["K"]), int(z["d"]), int(z["N"]), int(z["q"]), int(z["seed"])
    h, w = (int(v) for v in z["grid"])
    sm = torch.from_numpy(z["sum"]).cuda().unsqueeze(0).contiguous()
    cnt = torch.from_numpy(z["cnt"]).cuda().unsqueeze(0).contiguous()
    for name, t in (("hard", 3), ("gaussian", 0), ("gaussian", 40)):
        wts = z[f"som_{name}_{t}_w"]
        radius = (wts.shape[0] - 1) // 2
        wf = [float(v) for v in wts.reshape(-1)]
        osm, ocnt = torch.empty_like(sm), torch.empty_like(cnt)
        _lib.check(lib.rvq_som_spread(_p(sm), _p(cnt), _p(osm), _p(ocnt), (C.c_int * 2)(h, w), 1, K, d, radius,
                                      (C.c_float * len(wf))(*wf), _s()), "rvq_som_spread")
        assert np.array_equal(osm[0].cpu().numpy(), z[f"som_{name}_{t}_sum"])
        assert np.array_equal(ocnt[0].cpu().numpy(), z[f"som_{name}_{t}_cnt"])
    # re-seeding of stage q: stages before it have all-zero codebooks, so the stage-q residual is x itself
    nq = q + 1
    cb = torch.zeros(nq, K, d, device="cuda")
    x = torch.from_numpy(z["r_q"]).cuda().contiguous()
    idx = torch.zeros(N, nq, dtype=torch.int64, device="cuda")
    rep = torch.full((nq, K, d), 9.0, device="cuda")
    _lib.check(lib.rvq_reseed_gather(_p(x), N, N, 0, d, 1, d, nq, K, _p(cb), _p(idx), None, 0.99, 1.0, seed, 0, N,
                                     _p(rep), _s()), "rvq_reseed_gather")
    assert np.array_equal(rep[q].cpu().numpy(), z["rep"])
    half = x[N // 2:].contiguous()
    _lib.check(lib.rvq_reseed_gather(_p(half), N - N // 2, N - N // 2, 0, d, 1, d, nq, K, _p(cb), _p(idx[N // 2:].contiguous()),
                                     None, 0.99, 1.0, seed, N // 2, N, _p(rep), _s()), "rvq_reseed_gather")
    assert np.array_equal(rep[q].cpu().numpy(), z["rep_rank1"])
    cbq = torch.zeros(nq, K, d, device="cuda")
    ec = torch.full((nq, K), 7.0, device="cuda")
    es = torch.zeros(nq, K, d, device="cuda")
    cbq[q], ec[q], es[q] = torch.from_numpy(z["cb"]).cuda(), torch.from_numpy(z["ema_count"]).cuda(), torch.from_numpy(z["ema_sum"]).cuda()
    repf = torch.zeros(nq, K, d, device="cuda")
    repf[q] = torch.from_numpy(z["rep"]).cuda()
    nrep = torch.zeros(nq, dtype=torch.int32, device="cuda")
    _lib.check(lib.rvq_reseed_apply(_p(cbq), _p(ec), _p(es), _p(repf), None, nq, K, d, 1.0, 1.0, _p(nrep), _s()),
               "rvq_reseed_apply")
    assert nrep.tolist() == [0] * q + [int(z["n_replaced"])]
    assert np.array_equal(cbq[q].cpu().numpy(), z["new_cb"]) and np.array_equal(ec[q].cpu().numpy(), z["new_count"])
    assert np.array_equal(es[q].cpu().numpy(), z["new_sum"])
