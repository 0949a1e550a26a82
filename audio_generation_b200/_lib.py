"""ctypes binding of librvq_sm100a.so (C ABI declared in include/rvq_sm100a.h).

The library is the product: if it is missing or the device is not sm_100 every call raises.
There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "librvq_sm100a.so"

RVQ_ABI_VERSION = 7
RVQ_ALGO_TENSOR = 0
RVQ_ALGO_EXACT_SCAN = 1
# kernel selectors / launch options of rvq_encode's `flags` (include/rvq_sm100a.h)
RVQ_KERNELS = {"auto": 0x00, "generic": 0x10, "frame": 0x20, "tmem": 0x30}
RVQ_FLAG_CLUSTER_SHIFT = 8
RVQ_FLAG_COUNTERS = 0x1000

_vp, _i, _ll, _sz, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float

# name -> (restype, argtypes); must list every symbol include/rvq_sm100a.h declares
SIGNATURES = {
    "rvq_version": (_i, []),
    "rvq_last_error": (C.c_char_p, []),
    "rvq_device_supported": (_i, [_i]),
    "rvq_prepared_bytes": (_i, [_i, _i, _i, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "rvq_prepare_codebooks": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rvq_workspace_bytes": (_i, [_i, _i, _i, _ll, C.POINTER(_sz)]),
    "rvq_encode": (_i, [_vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                        _vp, _sz, _i, _vp]),
    "rvq_ema_finalize": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "rvq_ema_counts": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "rvq_dequantize": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _i, _i, C.POINTER(_f), _i, _vp, _vp]),
    "rvq_packed_bytes_per_frame": (_i, [_i, _i]),
    "rvq_pack_indices": (_i, [_vp, _ll, _i, _i, _vp, _vp]),
    "rvq_unpack_indices": (_i, [_vp, _ll, _i, _i, _vp, _vp]),
    "rvq_som_spread": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_i), _i, _i, _i, _i, C.POINTER(_f), _vp]),
    "rvq_reseed_frame": (C.c_ulonglong, [C.c_ulonglong, _i, _i, _i, C.c_ulonglong]),
    "rvq_reseed_gather": (_i, [_vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _i, _vp, _vp, _vp, _f, _f, C.c_ulonglong, _ll, _ll,
                               _vp, _vp]),
    "rvq_reseed_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp]),
    "rvq_backward": (_i, [_vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp]),
    "rvq_debug_stage_scores": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class RVQError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RVQError(
                f"{LIB_PATH} not found: build it with `python -m audio_generation_b200.build` "
                "(the RVQ path has no CPU/PyTorch fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rvq_last_error().decode(errors="replace")
        raise RVQError(f"{what} failed ({rc}): {msg}")
