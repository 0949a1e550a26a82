"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute without
a GPU), the drop-in module's construction / state / error behaviour, frame addressing."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from audio_generation_b200 import _lib, build
    build.build(verbose=False)
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "rvq_sm100a.h")).read()
    declared = set(re.findall(r"\b(rvq_[a-z_0-9]+)\s*\(", hdr)) - {"rvq_status"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rvq_version() == _lib.RVQ_ABI_VERSION == 7


def test_argument_checks_need_no_gpu():
    from audio_generation_b200 import _lib
    lib = _lib.load()
    ob, nb, mb = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.rvq_prepared_bytes(8, 1000, 128, ctypes.byref(ob), ctypes.byref(nb), ctypes.byref(mb)) == 0
    # norms: fp32 [nq, Kpad], their fp16 operand slices (32 bytes per code), the per-code allowance factor fp32
    # [nq, Kpad], its byte table [nq, Kpad] and one int flag per 256-code chunk (NormLayout, csrc/common.cuh)
    assert ob.value == 8 * 1024 * 128 * 2 and mb.value == 8 * 8 * 4
    assert nb.value == 8 * 1024 * (4 + 32 + 4 + 1) + 8 * (1024 // 256) * 4
    assert lib.rvq_prepared_bytes(0, 1, 1, None, None, None) == -1
    assert b"positive" in lib.rvq_last_error()
    n = ctypes.c_size_t()
    assert lib.rvq_workspace_bytes(8, 1024, 100, 10, ctypes.byref(n)) == -1      # d not a multiple of 64
    assert b"multiple" in lib.rvq_last_error()
    assert lib.rvq_workspace_bytes(8, 1024, 128, 10, ctypes.byref(n)) == 0 and n.value > 0


def test_module_is_a_drop_in_on_the_host_side():
    from audio_generation_b200 import ResidualQuantizer, tuple_checker
    import som_quantizer
    assert som_quantizer.ResidualQuantizer is ResidualQuantizer and som_quantizer.tuple_checker is tuple_checker
    m = ResidualQuantizer(num_quantizers=10, dim=512, quantizer_class="base", codebook_sizes=512,
                          vq_cutoff_freq=0.1, use_som=True, som_kernel_type="hard")     # config/training.yml kwargs
    assert m.num_quantizers == 10 and m.use_som and len(m.quantizers) == 10
    assert m.quantizers[0].som.height * m.quantizers[0].som.width == 512
    assert [n for n, _ in m.named_parameters()] == ["codebooks"]                         # "base": trainable codebooks
    e = ResidualQuantizer(8, 512, "ema", 1024)
    assert list(e.parameters()) == []                                                    # training.py:516 may be empty
    sd = e.state_dict()
    assert set(sd) == {"codebooks", "ema_count", "ema_sum", "k_valid", "update_steps"}                 # derived operands not persisted
    e2 = ResidualQuantizer(8, 512, "ema", 1024)
    e2.load_state_dict(sd)
    assert torch.equal(e2.codebooks, e.codebooks)
    assert len(e.get_stale_clusters()) == 8
    e.update_cutoff(ratio=0.95)
    r = ResidualQuantizer(3, 128, "ema", [512, 300, 64])
    assert r.K == 512 and r.k_valid.tolist() == [512, 300, 64]
    assert tuple_checker("a", 2) == ["a", "a"]


def test_no_cpu_fallback():
    from audio_generation_b200 import ResidualQuantizer
    from audio_generation_b200._lib import RVQError
    m = ResidualQuantizer(2, 64, "ema", 64)
    with pytest.raises(RVQError):
        m(torch.randn(4, 64))
    with pytest.raises(NotImplementedError):
        m(torch.randn(4, 64), prioritize_early=True)
    with pytest.raises(ValueError):
        m(torch.randn(4, 64), 3)


def test_frame_addressing():
    from audio_generation_b200.quantizer import _frame_addressing
    xc = torch.randn(3, 512, 150)
    xv = xc.permute(0, 2, 1)                     # reference layout: (B, L, d) view of (B, d, L)
    x3, N, L, sb, sl, sd = _frame_addressing(xv)
    assert x3.data_ptr() == xc.data_ptr() and (N, L, sb, sl, sd) == (450, 150, 512 * 150, 1, 150)
    x2 = torch.randn(100, 128)
    x3, N, L, sb, sl, sd = _frame_addressing(x2)
    assert x3.data_ptr() == x2.data_ptr() and (N, L, sl, sd) == (100, 100, 128, 1)
    odd = torch.randn(4, 10, 130)[:, :, 1:129]   # misaligned rows -> one contiguous copy
    x3, N, L, sb, sl, sd = _frame_addressing(odd)
    assert x3.is_contiguous() and (N, L, sl, sd) == (40, 10, 128, 1)


def test_product_path_never_imports_the_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import audio_generation_b200, som_quantizer; "
            "assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported by the product path'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "audio_generation_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_side_of_codebook_maintenance_matches_the_oracle():
    """The weights the module hands to rvq_som_spread (float32 after ctypes conversion) and the per-step seed of the
    re-seeding are the oracle's, for both neighbourhood kernels over the whole annealing range."""
    import ctypes as C
    import numpy as np
    from audio_generation_b200 import ResidualQuantizer
    from audio_generation_b200 import quantizer as Q
    from oracle import rvq_oracle as O
    for kernel in ("hard", "gaussian"):
        for t in (0, 1, 2, 7, 19, 20, 21, 100, 5000):
            r, w = Q.som_weights(kernel, t)
            ro, wo = O.som_weights(kernel, t)
            assert r == ro and len(w) == (2 * r + 1) ** 2
            as_f32 = np.array(list((C.c_float * len(w))(*w)), dtype=np.float32).reshape(2 * r + 1, 2 * r + 1)
            assert np.array_equal(as_f32, wo), (kernel, t)
    with pytest.raises(ValueError):
        ResidualQuantizer(2, 64, "ema", 64, som_kernel_type="soft")
    assert Q.approximate_square_root(1024) == O.approximate_square_root(1024) == (32, 32)
    assert Q.approximate_square_root(300) == O.approximate_square_root(300) == (15, 20)
    m = ResidualQuantizer(2, 64, "ema", 64, reseed_seed=77)
    ref = O.ResidualQuantizerRef(2, 64, "ema", 64, reseed_seed=77)
    for steps in (0, 1, 12345):
        ref.update_steps.fill_(steps)
        assert ((m.reseed_seed + steps * 0xD1B54A32D192ED03) & Q._M64) == ref.step_seed()
    # absolute-count staleness on the host side (no GPU needed): ema_count starts at 1, cutoff 1 -> nothing stale
    assert m.get_stale_clusters() == [0, 0]
    m.ema_count[1, :5] = 0.5
    assert m.get_stale_clusters() == [0, 5]
    m.update_cutoff(new_cutoff=0.25)
    assert m.get_stale_clusters() == [0, 0]


def test_module_survives_pickle_and_deepcopy():
    """torch.save(model) / copy.deepcopy keep the stage views bound to the copy and the load hook working."""
    import copy
    import io
    from audio_generation_b200 import ResidualQuantizer
    m = ResidualQuantizer(2, 64, "ema", [64, 48])
    m.update_steps.fill_(9)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    m3 = copy.deepcopy(m)
    for c in (m2, m3):
        assert c.quantizers[1]._p is c and c.codebook_sizes == [64, 48]
        c._steps_host = 3
        c.load_state_dict(m.state_dict())
        assert c._steps_host is None and int(c.update_steps) == 9


def test_clock_sampler_window_picks_samples_in_or_nearest_to_the_region():
    """bench.ClockSampler.window: samples inside the timed region (with its tolerance) when there are any, else the
    samples nearest to it - never an empty answer once the sampler has produced a row (an 8-GPU bench line once came
    back without clocks because the first sample arrived after a 75 ms region)."""
    import bench

    s = bench.ClockSampler()
    s.proc = object()                      # "running": window() only reads self.rows
    row = "{g}, {mhz}, 1965, 300.0, Not Active, Not Active, Not Active, {cap}"
    s.rows = [(100.00, row.format(g=0, mhz=1000, cap="Not Active")), (100.00, row.format(g=1, mhz=1100, cap="Not Active")),
              (100.50, row.format(g=0, mhz=1900, cap="Active")), (100.50, row.format(g=1, mhz=1800, cap="Not Active")),
              (101.00, row.format(g=0, mhz=1200, cap="Not Active")), (101.00, row.format(g=1, mhz=1300, cap="Not Active"))]
    inside = s.window(100.45, 100.55, [0, 1])
    assert inside[0]["sm_mhz"] == 1900 and inside[0]["reasons"] == ["sw_power_cap"] and inside[1]["sm_mhz"] == 1800
    nearest = s.window(100.70, 100.75, [0, 1])          # no sample inside: the closest ones (t = 100.5) are used
    assert nearest[0]["sm_mhz"] == 1900 and nearest[0]["samples"] == 1
    assert s.window(100.7, 100.75, [2]) == [None]        # a GPU the sampler never saw
    s.rows = []
    assert s.window(0.0, 1.0, [0]) == [None]
