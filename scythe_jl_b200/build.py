"""Build libscythe_b200.so (sm_100a) in-tree with nvcc.  `python -m scythe_jl_b200.build`.

Also builds the TEST-ONLY CPU emulation of the same kernel sources (tests/_emu/) when asked
with ``emu=True`` -- that library is never loaded by the product package.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
REPO = ROOT.parent
LIB = ROOT / "libscythe_b200.so"
EMU_DIR = REPO / "tests" / "_emu"
EMU_LIB = EMU_DIR / "libscythe_b200_emu.so"
SOURCES = ["sb_transforms.cu", "sb_ringfft.cu", "sb_ringfft2.cu", "sb_ringfft4.cu", "sb_chebmma.cu", "sb_model.cu", "sb_api.cpp", "sb_tables.cpp"]
HEADERS = ["sb_internal.hpp", "sb_fftcore.hpp", "sb_eqcore.hpp", "cuda_emu.h", "../../include/scythe_b200.h"]

NVCC_FLAGS = ["-O3", *os.environ.get("SB_NVCC_EXTRA", "").split(), "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    deps = [CSRC / s for s in SOURCES] + [CSRC / h for h in HEADERS]
    if not force and _newer(LIB, deps):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = ROOT / "build"
    objdir.mkdir(exist_ok=True)

    def compile_one(src):
        obj = objdir / (src + ".o")
        if not force and _newer(obj, [CSRC / src] + [CSRC / h for h in HEADERS]):
            return obj
        extra = ["-Xptxas", "-v"] if verbose else []
        xcu = ["-x", "cu"] if src.endswith(".cu") else []
        out = _run([nvcc, *NVCC_FLAGS, *extra, *xcu, "-c", str(CSRC / src), "-o", str(obj)])
        if verbose:
            print(out)
        return obj

    with ThreadPoolExecutor(4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_name(LIB.name + ".tmp")      # link beside the target, then rename: a reader never sees a partial library
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs),
          "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic", "-ldl", "-lcudart"])
    os.replace(tmp, LIB)
    return LIB


def build_emu(force: bool = False) -> Path:
    """TEST-ONLY: same kernel sources compiled for the host against csrc/cuda_emu.h."""
    srcs = SOURCES + ["cuda_emu.cpp"]
    deps = [CSRC / s for s in srcs] + [CSRC / h for h in HEADERS]
    if not force and _newer(EMU_LIB, deps):
        return EMU_LIB
    EMU_DIR.mkdir(parents=True, exist_ok=True)
    objdir = EMU_DIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src):
        obj = objdir / (src + ".o")
        if not force and _newer(obj, [CSRC / src] + [CSRC / h for h in HEADERS]):
            return obj
        _run(["g++", "-O2", "-g", "-std=c++20", "-DSB_EMU", "-fPIC", "-pthread", "-x", "c++", "-c", str(CSRC / src),
              "-o", str(obj)])
        return obj

    with ThreadPoolExecutor(5) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = EMU_LIB.with_name(EMU_LIB.name + ".tmp")
    _run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", "-o", str(tmp), *map(str, objs), "-ldl"])
    os.replace(tmp, EMU_LIB)
    return EMU_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    if "--emu" in sys.argv:
        print(build_emu(force))
    else:
        print(build(force, verbose="-v" in sys.argv))
