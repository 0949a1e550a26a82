"""CPU tests of the oracle (test infrastructure): golden vectors, an independent cross-check, the contract
the reference's call sites pin.  PARITY UNPINNED w.r.t. the upstream `som_quantizer` package."""
import os

import numpy as np
import pytest
import torch

from oracle import rvq_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_small():
    z = np.load(os.path.join(GOLD, "rvq_small.npz"))
    x, cbs = torch.from_numpy(z["x"]), torch.from_numpy(z["codebooks"])
    idx, xq, r, commit = O.rvq_encode_ref(x, list(cbs))
    assert np.array_equal(idx.numpy(), z["idx"])
    assert np.allclose(xq.numpy(), z["xq"], atol=1e-6)
    assert np.allclose(r.numpy(), z["resid"], atol=1e-6)
    assert np.allclose(np.array(commit), z["commit"], rtol=1e-6)
    res = O.stage_residuals_from_indices(x, list(cbs), idx)
    for q in range(cbs.shape[0]):
        cnt, sm = O.ema_stats_ref(res[q], idx[:, q], cbs.shape[1])
        ncb, nc, ns = O.ema_finalize_ref(cbs[q], torch.ones(cbs.shape[1]), cbs[q].clone(), cnt, sm)
        assert np.array_equal(cnt.numpy(), z["cnt"][q])
        assert np.allclose(sm.numpy(), z["sum"][q], atol=1e-5)
        assert np.allclose(ncb.numpy(), z["new_codebooks"][q], rtol=1e-5, atol=1e-6)


def _c1():
    z = np.load(os.path.join(GOLD, "c1_reference.npz"))
    nq, K, d = [int(v) for v in z["ctor"]]
    torch.manual_seed(int(z["codebook_seed"]))
    cbs = torch.randn(nq, K, d) * float(z["codebook_scale"]) * torch.tensor([0.8 ** i for i in range(nq)])[:, None, None]
    if abs(float(cbs.double().abs().sum()) - float(z["codebook_checksum"])) > 1e-6 * float(z["codebook_checksum"]):
        pytest.skip("torch CPU RNG stream differs from the one that generated the fixture")
    return z, cbs


def test_golden_c1_reference_latents():
    """BASELINE configs[0]: latents produced by the UNMODIFIED reference encoder on networks/om.wav."""
    z, cbs = _c1()
    assert tuple(z["x_shape"]) == (1, 136, 512) and tuple(z["x_stride"]) == (69632, 1, 136)   # vae.py:313 view
    assert tuple(z["ref_index_shape"]) == (1, 136, 10)
    x = torch.from_numpy(z["x_fp16"]).float().reshape(-1, 512)
    idx, xq, r, commit = O.rvq_encode_ref(x, list(cbs))
    assert np.array_equal(idx.numpy().astype(np.int16), z["idx"])
    assert np.allclose(np.array(commit), z["commit"], rtol=1e-5)
    assert abs(float(xq.double().sum()) - float(z["xq_checksum"])) < 1e-3 * max(1.0, abs(float(z["xq_checksum"])))


def test_cross_check_hf_encodec():
    """Independent implementation of the same chain: HF EncodecResidualVectorQuantizer.encode
    (distance -> max(-dist) -> residual subtract).  Same indices except fp32 near-ties."""
    tr = pytest.importorskip("transformers")
    from transformers import EncodecConfig
    from transformers.models.encodec.modeling_encodec import EncodecResidualVectorQuantizer
    torch.manual_seed(3)
    nq, K, d, L = 4, 128, 32, 300
    cfg = EncodecConfig(codebook_size=K, codebook_dim=d, hidden_size=d, target_bandwidths=[24.0], sampling_rate=24000)
    rvq = EncodecResidualVectorQuantizer(cfg)
    nq = min(nq, len(rvq.layers))
    cbs = []
    for q in range(nq):
        cb = torch.randn(K, d) * 0.7 ** q
        rvq.layers[q].codebook.embed.data.copy_(cb)
        cbs.append(cb)
    x = torch.randn(1, d, L)
    with torch.no_grad():
        codes = rvq.encode(x, bandwidth=None) if False else None
    # call the layers directly (bandwidth bookkeeping is irrelevant here)
    residual = x
    hf = []
    with torch.no_grad():
        for q in range(nq):
            i = rvq.layers[q].encode(residual)
            residual = residual - rvq.layers[q].decode(i)
            hf.append(i)
    hf = torch.stack(hf, -1).reshape(-1, nq)
    idx, _, _, _ = O.rvq_encode_ref(x[0].t().contiguous(), cbs, nq)
    adj = O.adjudicate_indices(x[0].t().contiguous(), cbs, hf)
    assert adj["n_illegal"] == 0
    assert (idx == hf).float().mean() > 0.995


def test_module_contract_cpu():
    """What the reference's call sites need."""
    torch.manual_seed(0)
    m = O.ResidualQuantizerRef(num_quantizers=3, dim=16, quantizer_class="ema", codebook_sizes=64,
                               vq_cutoff_freq=0.1, use_som=True, som_kernel_type="hard")
    xc = torch.randn(2, 16, 20)
    x = xc.permute(0, 2, 1)                       # the non-contiguous view of vae.py:313
    xq, idx, commit = m(x, 2, update_codebook=True)
    assert xq.shape == x.shape and idx.shape == (2, 20, 2) and idx.dtype == torch.int64 and commit.dim() == 0
    assert m.quantizers[1].dequantize(idx[:1, :, 1]).shape == (1, 20, 16)
    assert m.quantizers[0].som.height * m.quantizers[0].som.width == 64
    assert len(m.get_stale_clusters()) == 3
    m.update_cutoff(ratio=0.5)
    assert abs(m.vq_cutoff_freq - 0.05) < 1e-12
    with pytest.raises(NotImplementedError):
        m(x, prioritize_early=True)
    assert O.tuple_checker(5, 3) == [5, 5, 5] and O.tuple_checker((1, 2), 2) == (1, 2)
    with pytest.raises(AssertionError):
        O.tuple_checker((1, 2), 3)


def test_adjudicator_flags_real_errors():
    torch.manual_seed(1)
    cbs = [torch.randn(32, 8), torch.randn(32, 8) * 0.5]
    x = torch.randn(100, 8)
    idx, _, _, _ = O.rvq_encode_ref(x, cbs)
    assert O.adjudicate_indices(x, cbs, idx)["n_mismatch"] == 0
    bad = idx.clone()
    bad[5, 0] = (bad[5, 0] + 1) % 32
    a = O.adjudicate_indices(x, cbs, bad)
    assert a["n_mismatch"] >= 1 and a["n_illegal"] >= 1


def test_wire_format_oracle_known_answer():
    """Known answer of the LSB-first packing (10 bits per code, K = 1024) and the round trip."""
    import numpy as np
    from oracle import rvq_oracle as O
    idx = np.array([[1, 2, 3], [1023, 0, 512]], dtype=np.int64)
    packed = O.pack_indices_ref(idx, 10)
    # frame 0: 1 | 2 << 10 | 3 << 20 = 0x300801 -> bytes 01 08 30 00 ; frame 1: 0x3FF | 512 << 20 = 0x200003FF
    assert packed.tolist() == [[0x01, 0x08, 0x30, 0x00], [0xFF, 0x03, 0x00, 0x20]]
    assert (O.unpack_indices_ref(packed, 3, 10) == idx).all()
    rng = np.random.default_rng(0)
    for bits, n in [(9, 10), (10, 8), (12, 32), (1, 5), (16, 3)]:
        idx = rng.integers(0, 1 << bits, size=(50, n))
        assert (O.unpack_indices_ref(O.pack_indices_ref(idx, bits), n, bits) == idx).all()


def test_som_spread_oracle_known_answer():
    """Hand-computed 2 x 3 map, hard neighbourhood with sigma = 0.5: every cell = itself + 0.5 * its 4-neighbours."""
    import numpy as np
    radius, w = O.som_weights("hard", t=10, shrink=0.1)          # sigma = 1 / (1 + 1) = 0.5
    assert radius == 1 and w.tolist() == [[0, 0.5, 0], [0.5, 1, 0.5], [0, 0.5, 0]]
    cnt = np.array([1, 2, 4, 8, 16, 32, 99, 77], dtype=np.float32)       # K = 8: two padding codes beyond the map
    sm = np.stack([cnt, -cnt], axis=1)
    osm, ocnt = O.som_spread_ref(sm, cnt, 2, 3, radius, w)
    # grid [[1, 2, 4], [8, 16, 32]]
    assert ocnt.tolist() == [1 + 0.5 * (2 + 8), 2 + 0.5 * (1 + 4 + 16), 4 + 0.5 * (2 + 32),
                             8 + 0.5 * (1 + 16), 16 + 0.5 * (2 + 8 + 32), 32 + 0.5 * (4 + 16), 99, 77]
    assert (osm[:, 0] == ocnt).all() and (osm[:, 1] == -ocnt).all()
    # step 0: sigma = 1 -> plain 5-point sum; gaussian weights are symmetric and peak at the centre
    assert O.som_weights("hard", 0)[1][0, 1] == 1.0
    r, g = O.som_weights("gaussian", 0)
    assert r == 3 and g[3, 3] == 1.0 and (g == g.T).all() and (g == g[::-1, ::-1]).all()
    assert abs(float(g[3, 4]) - 0.60653066) < 1e-7
    with pytest.raises(ValueError):
        O.som_weights("soft", 0)


def test_reseed_oracle_matches_library_hash_and_known_answer():
    """The frame choice of the stale-code re-seeding is a stated contract (include/rvq_sm100a.h): the oracle's
    Python restatement, the library's host function and a hand-evaluated splitmix64 value must agree."""
    from audio_generation_b200 import _lib
    lib = _lib.load()
    # splitmix64 finaliser of 0x9E3779B97F4A7C15 (seed 0, q = k = 0) is the first output of SplitMix64(0)
    assert O.reseed_frame_ref(0, 0, 1024, 0, 1 << 64) == 0xE220A8397B1DCDAF
    for seed, q, K, k, total in [(0, 0, 1024, 0, 1 << 20), (123456789, 3, 512, 17, 600), (2 ** 64 - 1, 11, 1024, 1023, 7),
                                 (0xD1B54A32D192ED03, 5, 4096, 4095, 1 << 24)]:
        assert lib.rvq_reseed_frame(seed, q, K, k, total) == O.reseed_frame_ref(seed, q, K, k, total)
    cb = torch.arange(8.0).reshape(4, 2)
    cnt = torch.tensor([5.0, 0.2, 1.0, 0.99])
    rep = torch.full((4, 2), -1.0)
    ncb, nc, ns, n = O.reseed_apply_ref(cb, cnt, cb * 2, rep, cutoff=1.0, reset_count=1.0)
    assert n == 2 and nc.tolist() == [5.0, 1.0, 1.0, 1.0]
    assert ncb.tolist() == [[0, 1], [-1, -1], [4, 5], [-1, -1]] and ns[1].tolist() == [-1, -1] and ns[0].tolist() == [0, 2]


def test_module_update_with_som_and_reseed_cpu():
    """Ref module: a never-selected code goes stale and is replaced by a residual of the batch; replicas of the
    update (same seed) agree bit for bit; the SOM spread leaves the total count unchanged only for sigma -> 0."""
    torch.manual_seed(3)
    m = O.ResidualQuantizerRef(2, 8, "ema", 16, vq_cutoff_freq=1.0, use_som=True).train()
    with torch.no_grad():
        m.codebooks[0, 5] = 1e3                                   # far away: never selected
        m.ema_sum.copy_(m.codebooks)
        m.ema_count[0, 5] = 0.5
    x = torch.randn(64, 8)
    assert m.get_stale_clusters() == [1, 0]
    with torch.no_grad():
        m(x, None, update_codebook=True)
    assert int(m.update_steps) == 1
    # code (0, 5): count 0.5 * 0.99 + spread of its neighbours' hits; re-seeded only if still below the cutoff
    n = O.reseed_frame_ref(m.reseed_seed, 0, 16, 5, 64)
    assert float(m.ema_count[0, 5]) == 1.0 and torch.equal(m.codebooks[0, 5], x[n])      # stage 0 residual = x
    assert torch.equal(m.ema_sum[0, 5], x[n])


def test_oracle_reproduces_codebook_maintenance_fixture():
    """tests/golden/codebook_maint.npz (made by tests/golden/make_golden.py:maint) pins the SOM spreading, the frame
    choice of the re-seeding and the re-seeded state."""
    import numpy as np
    z = np.load(os.path.join(GOLD, "codebook_maint.npz"))
    K, d, N, q, seed = int(z["K"]), int(z["d"]), int(z["N"]), int(z["q"]), int(z["seed"])
    h, w = (int(v) for v in z["grid"])
    assert (h, w) == O.approximate_square_root(K)
    for name, t in (("hard", 3), ("gaussian", 0), ("gaussian", 40)):
        radius, wts = O.som_weights(name, t)
        assert np.array_equal(np.asarray(wts), z[f"som_{name}_{t}_w"])
        osm, ocnt = O.som_spread_ref(z["sum"], z["cnt"], h, w, radius, wts)
        assert np.array_equal(osm, z[f"som_{name}_{t}_sum"]) and np.array_equal(ocnt, z[f"som_{name}_{t}_cnt"])
    assert [O.reseed_frame_ref(seed, q, K, k, N) for k in range(K)] == z["frames"].tolist()
    r_q = torch.from_numpy(z["r_q"])
    rep = O.reseed_vectors_ref(r_q, q, K, seed)
    assert np.array_equal(rep.numpy(), z["rep"]) and np.array_equal(rep.numpy(), z["r_q"][z["frames"]])
    # a rank holding the second half of the frames contributes exactly the vectors it owns, zeros elsewhere
    r1 = O.reseed_vectors_ref(r_q[N // 2:], q, K, seed, frame_offset=N // 2, frames_total=N)
    assert np.array_equal(r1.numpy(), z["rep_rank1"])
    own = z["frames"] >= N // 2
    assert np.array_equal(r1.numpy()[own], z["rep"][own]) and not r1.numpy()[~own].any()
    ncb, nc, ns, n = O.reseed_apply_ref(torch.from_numpy(z["cb"]), torch.from_numpy(z["ema_count"]),
                                        torch.from_numpy(z["ema_sum"]), rep, 1.0, 1.0)
    assert n == int(z["n_replaced"]) == int((z["ema_count"] < 1.0).sum())
    assert np.array_equal(ncb.numpy(), z["new_cb"]) and np.array_equal(nc.numpy(), z["new_count"])
    assert np.array_equal(ns.numpy(), z["new_sum"])


def test_codebook_maintenance_properties():
    """Size-independent properties of the maintenance restatement: the neighbourhood is linear, reduces to the identity
    for a centre-only kernel, conserves mass away from the border for the hard kernel; the frame choice of the
    re-seeding is deterministic, in range and spread over the batch."""
    import numpy as np
    rng = np.random.default_rng(4)
    h, w, d = 6, 9, 5
    a, b = rng.standard_normal((h * w, d)).astype(np.float32), rng.standard_normal((h * w, d)).astype(np.float32)
    ca, cb_ = rng.integers(0, 7, h * w).astype(np.float32), rng.integers(0, 7, h * w).astype(np.float32)
    r, wt = O.som_weights("gaussian", 5)
    sa, na = O.som_spread_ref(a, ca, h, w, r, wt)
    sb, nb = O.som_spread_ref(b, cb_, h, w, r, wt)
    sab, nab = O.som_spread_ref(a + b, ca + cb_, h, w, r, wt)
    assert np.allclose(sab, sa + sb, atol=1e-5) and np.allclose(nab, na + nb, atol=1e-5)
    ident = np.zeros((3, 3), dtype=np.float32)
    ident[1, 1] = 1.0
    s1, n1 = O.som_spread_ref(a, ca, h, w, 1, ident)
    assert np.array_equal(s1, a) and np.array_equal(n1, ca)
    # hard kernel, sigma = 0.25: a unit count in an interior cell spreads to 1 + 4 * 0.25 = 2 in total
    one = np.zeros(h * w, dtype=np.float32)
    one[3 * w + 4] = 1.0
    _, n2 = O.som_spread_ref(np.zeros((h * w, d), np.float32), one, h, w, *O.som_weights("hard", 30))
    assert abs(float(n2.sum()) - 2.0) < 1e-6 and int((n2 > 0).sum()) == 5
    K, N = 1024, 600
    fr = [O.reseed_frame_ref(99, 3, K, k, N) for k in range(K)]
    assert fr == [O.reseed_frame_ref(99, 3, K, k, N) for k in range(K)] and 0 <= min(fr) and max(fr) < N
    assert len(set(fr)) > 0.7 * N                      # 1024 draws over 600 frames hit most of them
    assert fr != [O.reseed_frame_ref(100, 3, K, k, N) for k in range(K)]
    assert fr != [O.reseed_frame_ref(99, 4, K, k, N) for k in range(K)]


def test_base_class_keeps_usage_counts_reseeds_and_spreads_gradient():
    """quantizer_class="base" with update_codebook=True (what config/training.yml + training.py:305-308 ask for):
    usage counts are averaged, stale codes re-seeded, the codebooks otherwise only move by gradient, and with use_som
    the gradient of a winning code reaches its four map neighbours with weight sigma_t."""
    torch.manual_seed(0)
    nq, K, d = 2, 16, 8
    m = O.ResidualQuantizerRef(nq, d, "base", K, vq_cutoff_freq=0.0, use_som=True, som_kernel_type="hard")
    m.train()
    cb0 = m.codebooks.detach().clone()
    x = torch.randn(1, 40, d)
    _, idx, commit = m(x, None, update_codebook=True)
    commit.backward()
    assert torch.equal(m.codebooks.detach(), cb0)                        # no EMA refresh for "base"
    cnt = torch.bincount(idx[0, :, 0], minlength=K).float()
    assert torch.allclose(m.ema_count[0], 0.99 * torch.ones(K) + 0.01 * cnt)
    assert int(m.update_steps) == 1
    # gradient spreading: a code nobody selected still receives gradient if a grid neighbour was selected
    h, w = O.approximate_square_root(K)
    g = m.codebooks.grad[0]
    hit = cnt > 0
    for k in range(K):
        y, xg = divmod(k, w)
        nb = [(y + dy) * w + (xg + dx) for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1))
              if 0 <= y + dy < h and 0 <= xg + dx < w]
        reached = bool(hit[k]) or any(bool(hit[n]) for n in nb)
        assert (g[k].abs().sum() > 0) == reached
    # re-seeding: with a cutoff above every count all codes are replaced by residual rows
    m2 = O.ResidualQuantizerRef(nq, d, "base", K, vq_cutoff_freq=5.0, use_som=False)
    m2.train()
    m2(x, None, update_codebook=True)
    assert m2.n_replaced == [K, K] and m2.get_stale_clusters() == [0, 0]
    rows = x.reshape(-1, d)
    assert all(any(torch.equal(m2.codebooks[0, k].detach(), r) for r in rows) for k in range(K))
