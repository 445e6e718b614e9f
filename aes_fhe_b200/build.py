"""In-tree builds.  `build_cuda()` compiles the product library for sm_100a with nvcc
(cross-compiles without a GPU).  `build_emu()` is used by tests only: the same sources
compiled by g++ with -DFHE_EMU into tests/emu/ (see csrc/compat.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
SOURCES = ["capi.cu"]
HEADERS = ["compat.h", "modarith.cuh", "ntt.cuh", "ntt_chained.cuh", "ntt_fused.cuh", "kernels.cuh"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-DFHE_PASSA_MINB=4", "-DFHE_PASSB_MINB=4",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps if Path(d).exists())


def _deps():
    return [CSRC / f for f in SOURCES + HEADERS] + [ROOT / "include" / "aesfhe_b200.h"]


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    out = CSRC / "libaesfhe_b200.so"
    if not force and not _stale(out, _deps()):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-ccbin", "g++", "-o", str(out)] + \
          [str(CSRC / s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=str(CSRC))
    return out


def build_profile_variant() -> Path:
    """csrc/variants/libprof.so: the product sources with -DFHE_FUSED_PROFILE (cycle counters inside the
    persistent fused NTT), used by tools/fused_prof.py only."""
    out = CSRC / "variants" / "libprof.so"
    out.parent.mkdir(exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DFHE_FUSED_PROFILE", "-ccbin", "g++", "-o", str(out)] +
                          [str(CSRC / s) for s in SOURCES], cwd=str(CSRC))
    return out


def build_emu(force: bool = False) -> Path:
    emu_dir = ROOT / "tests" / "emu"
    out = emu_dir / "libaesfhe_emu.so"
    deps = _deps() + [emu_dir / "emu_runtime.cpp"]
    if not force and not _stale(out, deps):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DFHE_EMU", "-ffp-contract=off", "-pthread", f"-I{CSRC}",
           "-x", "c++"] + [str(CSRC / s) for s in SOURCES] + [str(emu_dir / "emu_runtime.cpp"), "-o", str(out)]
    subprocess.check_call(cmd)
    return out
