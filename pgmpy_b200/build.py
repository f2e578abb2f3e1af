"""Builds libpgx.so (sm_100a) in-tree with nvcc. No JIT cache: the .so travels with the source tree."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgx.so")
SOURCES = [os.path.join(CSRC, "pgx.cu")]


def _headers():
    """Every header the library is compiled from: a stale .so must never be benchmarked silently."""
    import glob

    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))) + [
        os.path.join(os.path.dirname(HERE), "include", "pgx.h")]



def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + _headers())


def build_native(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC,-O3",
        "-shared",
        "-o", LIB,
    ] + SOURCES
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
