"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see amg_oracle.c header).

  python oracle/build.py          # or: from oracle.build import build_oracle; build_oracle()

Outputs (both git-ignored, both travel to the GPU box with the gpurun snapshot):
  oracle/libamg_oracle.so   the plain-C restatement (amg_oracle.c), gcc -O2, no fast-math
  oracle/_ref/libamg_ref.so the REFERENCE's own relaxation.h compiled from where it lies
                            under /root/reference (ref_shim.cpp only instantiates it);
                            built only when /root/reference exists (this container).
Flags mirror the reference's meson.build:4,7 (-O2, c++11).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CORE = "/root/reference/pyamg/amg_core"


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def build_oracle(force=False, verbose=False):
    """Compile the oracle .so files if missing or stale. Returns dict of paths (None if absent)."""
    out = {"oracle": None, "ref": None}
    src = os.path.join(HERE, "amg_oracle.c")
    lib = os.path.join(HERE, "libamg_oracle.so")
    if force or _stale(lib, [src]):
        cmd = ["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", lib, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    out["oracle"] = lib

    shim = os.path.join(HERE, "ref_shim.cpp")
    refdir = os.path.join(HERE, "_ref")
    reflib = os.path.join(refdir, "libamg_ref.so")
    if os.path.isdir(REF_CORE):
        os.makedirs(refdir, exist_ok=True)
        if force or _stale(reflib, [shim, os.path.join(REF_CORE, "relaxation.h")]):
            cmd = ["g++", "-O2", "-std=c++11", "-shared", "-fPIC", "-I", REF_CORE,
                   "-o", reflib, shim]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if os.path.exists(reflib):
        out["ref"] = reflib
    return out


if __name__ == "__main__":
    print(build_oracle(force=True, verbose=True))
