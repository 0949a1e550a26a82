"""Build librvq_sm100a.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m audio_generation_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = Path(__file__).resolve().parent / "librvq_sm100a.so"
SOURCES = ["rvq_abi.cu", "rvq_aux.cu", "rvq_encode_tc.cu", "rvq_encode_tr.cu", "rvq_encode_fr.cu", "rvq_codebook.cu", "rvq_backward.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "exact.cuh", "encode_common.cuh", "../../include/rvq_sm100a.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [Path(__file__)]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = True) -> Path:
    if not force and not needs_build():
        return LIB
    # extra -D switches for experiment builds (e.g. RVQ_NVCC_DEFS="-DRVQ_FR_PROFILE"): read at BUILD time only
    extra = os.environ.get("RVQ_NVCC_DEFS", "").split()
    objs = []
    log = []
    for src in SOURCES:
        obj = CSRC / (src[:-3] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(str(obj))
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    # register / spill report of every kernel (compile times dropped: they would change the file on every build)
    info = [ln for ln in "\n".join(log).splitlines() if "Compile time" not in ln]
    (CSRC / "ptxas_info.txt").write_text("\n".join(info) + "\n")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
