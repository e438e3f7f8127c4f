"""Build libndb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU.  One object per translation unit, compiled in parallel, cached by
source mtime; the final link is a plain `nvcc -shared`.  The library links cudart statically and has
no torch / Python dependency: it is the drop-in boundary (include/ndb200.h).
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
OBJ = ROOT / "build"
LIB = ROOT / "libndb200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # Bit-exact parity with Rust: never contract a*b+c into an FMA (kernels that want one call fma()).
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-DNDB_BUILDING",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libndb200 has no non-CUDA build")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "ndb200.h"]
    return max(p.stat().st_mtime for p in hdrs)


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _deps_mtime())
    if obj.exists() and obj.stat().st_mtime > newest:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr.strip():
        print(r.stderr, file=sys.stderr)
    return obj


def build(verbose: bool = False, jobs: int | None = None) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sources()
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if LIB.exists() and all(LIB.stat().st_mtime > o.stat().st_mtime for o in objs):
        return LIB
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
           "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
