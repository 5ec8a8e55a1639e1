"""In-tree build of librssm_rollout.so (nvcc, sm_100a only).  Works without a GPU (cross-compiles)."""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "librssm_rollout.so"
SOURCES = ("mrssm_kernels.cu", "mrssm_wide_fwd.cu", "mrssm_wide_bwd.cu", "mrssm_wide_wgrad.cu", "mtrssm_kernels.cu", "mtrssm_fwd2.cu", "mtrssm_fused_bwd.cu", "wgrad_kernel.cu", "likelihood_kernel.cu", "p2p_allreduce.cu", "rollout_abi.cu")
NVCC_FLAGS = (
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build librssm_rollout.so")


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "rssm_rollout.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_stamp() -> str:
    """Hash of the sources + flags the shipped library was built from (bench.py keys its ncu traffic figures by it)."""
    return _stamp()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA sources into `librssm_rollout.so` next to this file (skipped when up to date).

    One process per GPU means N ranks may get here at once: the stamp check and the build run under an exclusive `flock`, the
    objects go to a per-process directory and the library is linked to a temporary name and `os.replace`d into place, so a rank
    can never load a half-written file."""
    import fcntl

    obj_root = PKG / "build"
    obj_root.mkdir(exist_ok=True)
    with open(obj_root / "lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    stamp_file = PKG / "build" / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    obj_dir = PKG / "build"
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc, *flags, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs)]  # cudart linked statically (nvcc default)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB)
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
