"""In-tree build of libpfst_sm100.so (nvcc, sm_100a only).

``python -m pfst_b200.build`` or ``pfst_b200.build.build()``. The shared library
is written next to the sources (``pfst_b200/csrc/libpfst_sm100.so``) so that it
travels with a snapshot of the repo; it is git-ignored. nvcc cross-compiles
without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libpfst_sm100.so"
OBJ_DIR = CSRC / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; pfst_b200 has no non-CUDA fallback")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((CSRC.parent.parent / "include").glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs)


def _compile(src: Path, verbose: bool) -> tuple[Path, str]:
    obj = OBJ_DIR / (src.stem + ".o")
    newest_dep = max(src.stat().st_mtime, _deps_mtime())
    if obj.exists() and obj.stat().st_mtime >= newest_dep:
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("PFST_EXTRA_NVCC", "").split(), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(exist_ok=True)
    if force:
        for o in OBJ_DIR.glob("*.o"):
            o.unlink()
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    if (not LIB.exists()) or force or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(path)
