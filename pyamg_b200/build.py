"""In-tree build of libpyamg_b200.so (sm_100a only).

  python -m pyamg_b200.build            # or __graft_entry__.build()

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libpyamg_b200.so")
SOURCES = [os.path.join(PKG, "csrc", "engine.cu")]
HEADERS = [os.path.join(PKG, "csrc", f) for f in sorted(os.listdir(os.path.join(PKG, "csrc")))
           if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "pyamg_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _digest(paths):
    """Content hash of the sources: file times do not survive the copy to the GPU box."""
    import hashlib
    h = hashlib.sha256()
    for p in sorted(paths):
        if os.path.exists(p):
            h.update(os.path.basename(p).encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()


def _stamp_ok(lib, paths):
    stamp = lib + ".sha256"
    return os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == _digest(paths)


def _write_stamp(lib, paths):
    with open(lib + ".sha256", "w") as f:
        f.write(_digest(paths))


def is_stale():
    return not _stamp_ok(LIB, SOURCES + HEADERS)


HOST_LIB = os.path.join(PKG, "libpyamg_b200_host.so")
HOST_SOURCES = [os.path.join(PKG, "csrc", "host_setup.cpp")]


def build_host_library(force=False, verbose=False):
    """Compile the host-side setup helpers (plain C++, no CUDA). Returns the .so path."""
    if not force and _stamp_ok(HOST_LIB, HOST_SOURCES):
        return HOST_LIB
    cxx = shutil.which("g++")
    if cxx is None:
        if os.path.exists(HOST_LIB):
            return HOST_LIB
        raise RuntimeError("g++ not found and libpyamg_b200_host.so is not built")
    cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-o", HOST_LIB] + HOST_SOURCES
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    try:
        subprocess.check_call(cmd)
    except subprocess.CalledProcessError:          # toolchain without OpenMP: the pragmas are ignored
        cmd.remove("-fopenmp")
        subprocess.check_call(cmd)
    _write_stamp(HOST_LIB, HOST_SOURCES)
    return HOST_LIB


def build_extension(force=False, verbose=False):
    """Compile the CUDA engine if missing or older than its sources. Returns the .so path."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB   # GPU box without nvcc on PATH: use the prebuilt library
        raise RuntimeError("nvcc not found and libpyamg_b200.so is not built")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + SOURCES
    if os.environ.get("AMGB_PTXAS_V"):
        cmd.insert(1, "-Xptxas=-v")
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    _write_stamp(LIB, SOURCES + HEADERS)
    return LIB


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose=True))
    print(build_host_library(force="--force" in sys.argv, verbose=True))
