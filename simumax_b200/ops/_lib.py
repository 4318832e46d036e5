"""In-tree build + ctypes loader for the sm_100a kernel library.

All kernels live in ``simumax_b200/csrc/*.cu`` and are compiled by plain ``nvcc`` into one shared
object ``simumax_b200/_kernels_sm100a.so`` (git-ignored, but it travels to the GPU box with the
snapshot).  The Python side talks to it through a C ABI with ``ctypes`` -- raw device pointers and a
``cudaStream_t`` -- so there is no dependency on torch's C++ headers and a full rebuild takes seconds.

On a machine with a GPU the ops fail loudly if the library is missing; there is no silent PyTorch
fallback.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent.parent
CSRC = _PKG / "csrc"
LIB_PATH = _PKG / "_kernels_sm100a.so"
_STAMP = _PKG / "_kernels_sm100a.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
    "-Xcompiler", "-fPIC",
]

_lib = None


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _source_hash() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_built() -> bool:
    return LIB_PATH.exists() and _STAMP.exists() and _STAMP.read_text().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a into one .so (parallel per-file, then link)."""
    if not force and is_built():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = _PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append((src, out))
        elif verbose and out:
            print(out, file=sys.stderr)
    if failed:
        msg = "\n".join(f"--- {s} ---\n{o}" for s, o in failed)
        raise RuntimeError(f"nvcc failed for sm_100a kernels:\n{msg}")
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    _STAMP.write_text(_source_hash())
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (building if needed and possible) the kernel library.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    elif not is_built() and os.environ.get("SMX_NO_REBUILD", "0") != "1":
        try:
            build()
        except Exception as e:  # stale .so but no nvcc: keep going with what we have, loudly
            print(f"[simumax_b200] WARNING: kernel sources changed but rebuild failed: {e}", file=sys.stderr)
    _lib = ctypes.CDLL(str(LIB_PATH))
    return _lib


def available() -> bool:
    try:
        lib()
        return True
    except Exception:
        return False


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"simumax_b200 kernel '{what}' failed with code {rc}")


_LAUNCH_COUNT = 0


def count_launch(n: int = 1) -> None:
    """Every op wrapper reports its kernel launches here (bench.py's `gpu_launches`)."""
    global _LAUNCH_COUNT
    _LAUNCH_COUNT += n


def launches() -> int:
    return _LAUNCH_COUNT


def reset_launches() -> None:
    global _LAUNCH_COUNT
    _LAUNCH_COUNT = 0
