"""Builds libpgx.so (sm_100a) in-tree with nvcc. No JIT cache: the .so travels with the source tree."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgx.so")
SOURCES = [os.path.join(CSRC, "pgx.cu"), os.path.join(CSRC, "pgx_mm.cu"), os.path.join(CSRC, "pgx_tc32.cu"),
           os.path.join(CSRC, "pgx_spec.cu")]
OBJ_DIR = os.path.join(CSRC, "_obj")


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
    """One nvcc per translation unit, in parallel, then one link: sm_100a only, -lineinfo for ncu's source page."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    flags += [f"-D{d}" for d in os.environ.get("PGX_DEFINES", "").split() if d]  # tuning builds only
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        procs.append((obj, subprocess.Popen([nvcc_path(), *flags, "-c", src, "-o", obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.PIPE, text=True)))
    log = ""
    for obj, p in procs:
        so, se = p.communicate()
        log += so + se
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + so + se)
    res = subprocess.run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [o for o, _ in procs] + ["-ldl"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
