"""Build recipe of libperfb200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun).

    python -m perf_b200.build [--force]
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libperfb200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


class NvccMissing(RuntimeError):
    """No nvcc on this machine (the only build failure a caller may answer by loading the shipped .so)."""


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise NvccMissing("nvcc not found: libperfb200.so cannot be built on this machine")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


STAMP = os.path.join(HERE, ".libperfb200.srchash")


def _deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "perfb200.h")]


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    """True when libperfb200.so is missing or was built from other sources than the ones on disk: compared by content
    (the stamp file written by build() travels with the .so), never by modification time."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        return open(STAMP).read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into perf_b200/libperfb200.so.  Returns the path."""
    if not force and not is_stale():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:      # several ranks may import at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():                            # another process built it meanwhile
            return LIB
        tmp = f"{LIB}.{os.getpid()}.tmp"
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", tmp]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        os.replace(tmp, LIB)
        with open(STAMP + ".tmp", "w") as f:
            f.write(source_hash() + "\n")
        os.replace(STAMP + ".tmp", STAMP)
        if verbose:
            print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
