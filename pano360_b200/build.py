"""Builds ``libpano360_b200.so`` in-tree with nvcc for sm_100a.

The shared library is the product's only compute path; it is git-ignored but
travels to the GPU box with the snapshot.  ``python -m pano360_b200.build``
or ``__graft_entry__.build()`` runs this.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpano360_b200.so")
SOURCES = ["p360_api.cu", "p360_warp.cu", "p360_blend.cu", "p360_blur.cu", "p360_gain.cu", "p360_pyramid.cu", "p360_crop.cu", "p360_exact.cu", "p360_resize.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "shared"]


def nvcc():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def stale():
    """Missing, or older than a source.  A library that exists where nvcc does not (a snapshot
    copied to a box without the toolkit) is never declared stale: it cannot be rebuilt there and
    copying does not preserve mtimes reliably."""
    if not os.path.exists(LIB):
        return True
    if not os.path.exists(nvcc()) and not any(
            os.access(os.path.join(d, "nvcc"), os.X_OK) for d in os.environ.get("PATH", "").split(os.pathsep) if d):
        return False
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pano360_b200.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + os.environ.get("P360_NVCC_DEFS", "").split()      # (-D... for A/B builds)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    # One builder at a time (N ranks of a torchrun may all find the library stale), and the
    # library appears atomically: built next to it, then renamed over it.
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not stale():          # somebody else built it while we waited
            return LIB
        tmp = f"{LIB}.{os.getpid()}.tmp"
        full = cmd + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
        res = subprocess.run(full, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed:\n" + " ".join(full) + "\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
