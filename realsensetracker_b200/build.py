"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

  _lib/librst_align.so : CUDA kernels + C ABI (include/rst_align.h), nvcc, sm_100a only
  _lib/librst_synth.so : synthetic RGB-D frame source (CPU, gcc + OpenMP)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]
NVCC_LINK_FLAGS = ["--shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]

# /opt/gcc/bin (the image's $CC) lacks libgomp.spec; the system compilers have it.
HOST_CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(out: Path, srcs) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in srcs)


def _run(cmd, log_name=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log_name:
        (LIBDIR / log_name).write_text(" ".join(map(str, cmd)) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd))}")
    return res.stdout


def build_align(force: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "librst_align.so"
    cu = sorted(CSRC.glob("*.cu"))
    deps = cu + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inl")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "rst_align.h"]
    if force or _stale(out, deps):
        # one object per translation unit, compiled concurrently (no cross-file device calls), then one link
        from concurrent.futures import ThreadPoolExecutor
        objdir = LIBDIR / "obj"
        objdir.mkdir(exist_ok=True)
        hdrs = [d for d in deps if d not in cu]

        def compile_one(src: Path) -> str:
            obj = objdir / (src.stem + ".o")
            log = objdir / (src.stem + ".log")   # ptxas -v output of this translation unit (kept across incremental builds)
            if force or _stale(obj, [src, *hdrs]) or not log.exists():
                log.write_text(_run([_nvcc(), *NVCC_FLAGS, "-ccbin", HOST_CXX, "-c", "-o", str(obj), str(src)]))
            return log.read_text()

        with ThreadPoolExecutor(max_workers=len(cu)) as ex:
            logs = list(ex.map(compile_one, cu))
        (LIBDIR / "nvcc_ptxas.log").write_text("\n".join(logs))
        _run([_nvcc(), *NVCC_LINK_FLAGS, "-ccbin", HOST_CXX, "-o", str(out), *[str(objdir / (c.stem + ".o")) for c in cu]])
    return out


def build_synth(force: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "librst_synth.so"
    src = CSRC / "rst_synth.c"
    if force or _stale(out, [src, CSRC / "rst_synth.h"]):
        _run([HOST_CC, "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=c11",
              "-o", str(out), str(src), "-lm"])
    return out


def build_all(force: bool = False):
    return build_align(force), build_synth(force)


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print(p)
