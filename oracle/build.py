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




# ---------------------------------------------------------------------------------------------------------------
# The WHOLE reference package, importable (SURVEY 8(c) recipe): test infrastructure and the `--impl reference` arm.
# ---------------------------------------------------------------------------------------------------------------
REF_PKG = "/root/reference/pyamg"
# the contract's place for the unmodified reference: <repo>/baseline/_ref (git-ignored, NOT gpurun-ignored: it travels
# to the GPU box).  `pip install --target baseline/_ref /root/reference` needs meson-python, which this image lacks, so
# the same install is produced by hand below.
REF_SITE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
_EXT = ["air", "evolution_strength", "graph", "krylov", "linalg", "relaxation", "ruge_stuben", "smoothed_aggregation"]


def reference_site():
    """Path to put on sys.path so that `import pyamg` is the unmodified reference (None if it was never built)."""
    return REF_SITE if os.path.exists(os.path.join(REF_SITE, "pyamg", "__init__.py")) else None


def build_reference_package(force=False, verbose=False, jobs=8):
    """Install the unmodified reference into baseline/_ref: the package tree is copied there as an INSTALL (the
    directory is git-ignored; nothing of it enters this repo's history) and its eight checked-in pybind11 binding
    units are compiled in place with the flags of the reference's meson.build:4,7 -- meson itself is not in this
    image.  A dist-info stub answers pyamg/__init__.py:12-13's importlib.metadata.version call."""
    import shutil
    import sysconfig
    if not os.path.isdir(REF_PKG):
        return reference_site()
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    dst = os.path.join(REF_SITE, "pyamg")
    done = all(os.path.exists(os.path.join(dst, "amg_core", h + suffix)) for h in _EXT)
    if done and not force:
        return REF_SITE
    if os.path.isdir(REF_SITE):
        shutil.rmtree(REF_SITE)
    os.makedirs(REF_SITE)
    shutil.copytree(REF_PKG, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    inc = subprocess.check_output(["python3", "-m", "pybind11", "--includes"], text=True).split()
    core = os.path.join(dst, "amg_core")
    procs = []
    for h in _EXT:
        cmd = ["g++", "-O2", "-std=c++11", "-ftemplate-depth=2048", "-shared", "-fPIC", "-fvisibility=hidden",
               *inc, h + "_bind.cpp", "-o", h + suffix]
        if verbose:
            print(" ".join(cmd))
        procs.append(subprocess.Popen(cmd, cwd=core))
        if len(procs) >= jobs:
            for p in procs:
                assert p.wait() == 0
            procs = []
    for p in procs:
        assert p.wait() == 0
    info = os.path.join(REF_SITE, "pyamg-0.0.0.dist-info")
    os.makedirs(info, exist_ok=True)
    with open(os.path.join(info, "METADATA"), "w") as f:
        f.write("Metadata-Version: 2.1\nName: pyamg\nVersion: 0.0.0+oracle\n")
    return REF_SITE


def import_reference():
    """`import pyamg` = the unmodified reference from baseline/_ref (raises ImportError when it was never built).
    Only tests/, smoke() and bench.py's reference / cpu_baseline legs may call this."""
    import sys
    site = reference_site()
    if site is None:
        raise ImportError("baseline/_ref is absent: run oracle/build.py where /root/reference exists")
    if site not in sys.path:
        sys.path.insert(0, site)
    import pyamg
    return pyamg


if __name__ == "__main__":
    print(build_oracle(force=True, verbose=True))
    print(build_reference_package(force=True, verbose=True))
