"""Build tests/emu/_gen/libpyamg_b200_emu.so: the engine's own sources (pyamg_b200/csrc/*.cu, *.cuh) compiled
with g++ against tests/emu/cuda_runtime.h, so that the SAME kernel code runs on the CPU of the GPU-less build
container (logic validation only -- see the header of cuda_runtime.h).  TEST INFRASTRUCTURE: nothing under
pyamg_b200/ knows this library exists; tests/conftest.py loads it when AMGB_TEST_EMU=1.

Two purely syntactic rewrites are applied to a copy of the sources (line numbers are preserved, `#line`
directives point the compiler and gdb at the originals):
  kernel<<<grid, block, smem, stream>>>(args)   ->  ::emu::launch_kernel(::emu::LaunchCfg(grid, block, smem, stream), (kernel), args)
  extern __shared__ ... name[];                 ->  unsigned char *name = ::emu::dyn_smem();
The inline-PTX wrappers carry their own `#ifdef AMGB_EMU` bodies in the product sources.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pyamg_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
# AMGB_EMU_SANITIZE=1: a second library built with AddressSanitizer (device allocations are host heap blocks, so an
# out-of-bounds global-memory access in a kernel is reported).  Run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# and ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0.
SANITIZE = os.environ.get("AMGB_EMU_SANITIZE") == "1"
LIB = os.path.join(GEN, "libpyamg_b200_emu_asan.so" if SANITIZE else "libpyamg_b200_emu.so")


def _match_back_template(s, i):
    """s[i] == '>': index of the matching '<' scanning backwards."""
    depth = 0
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template brackets before <<<")


def _match_paren(s, i):
    """s[i] == '(': index of the matching ')'."""
    depth = 0
    while i < len(s):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses after >>>")


def rewrite_launches(src):
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier (with ::) and an optional template argument list, ending at k
        e = k
        while src[e - 1].isspace():
            e -= 1
        b = e
        if src[b - 1] == ">":
            b = _match_back_template(src, b - 1)
        while b > 0 and (src[b - 1].isalnum() or src[b - 1] in "_:"):
            b -= 1
        kernel = src[b:e]
        c1 = src.find(">>>", k)
        cfg = src[k + 3:c1]
        a0 = c1 + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == "(", "expected an argument list after >>>: " + src[k - 40:k + 80]
        a1 = _match_paren(src, a0)
        args = src[a0 + 1:a1]
        out.append(src[pos:b])
        out.append("::emu::launch_kernel(::emu::LaunchCfg(%s), (%s)%s%s)" % (cfg, kernel, ", " if args.strip() else "", args))
        pos = a1 + 1


_EXTERN_SHARED = re.compile(r"extern\s+__shared__[^;]*?\b(\w+)\s*\[\s*\]\s*;")


def transform(text, origin):
    text = rewrite_launches(text)
    text = _EXTERN_SHARED.sub(lambda m: "unsigned char *%s = ::emu::dyn_smem();" % m.group(1), text)
    return '#line 1 "%s"\n%s' % (origin, text)


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files += [os.path.join(HERE, "cuda_runtime.h"), os.path.join(HERE, "build_emu.py")]
    for p in files:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    h.update(open(os.path.join(ROOT, "include", "pyamg_b200.h"), "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Returns the path of the emulation library (rebuilt when any source changed)."""
    import fcntl
    stamp = LIB + ".sha256"
    dig = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig
    if not force and fresh():
        return LIB
    os.makedirs(GEN, exist_ok=True)
    with open(os.path.join(GEN, ".lock"), "w") as lock:       # several ranks / test processes may get here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return LIB
        return _build_locked(dig, stamp, verbose)


def _build_locked(dig, stamp, verbose):
    gdir = os.path.join(GEN, "pyamg_b200", "csrc")
    os.makedirs(gdir, exist_ok=True)
    os.makedirs(os.path.join(GEN, "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "pyamg_b200.h"), os.path.join(GEN, "include", "pyamg_b200.h"))
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            src = open(os.path.join(CSRC, f)).read()
            with open(os.path.join(gdir, f), "w") as o:
                o.write(transform(src, os.path.join(CSRC, f)))
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-DAMGB_EMU", "-I", HERE, "-fPIC", "-shared",
           "-fno-strict-aliasing", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
           os.path.join(gdir, "engine.cu"), "-o", LIB + ".tmp"]
    if SANITIZE:
        cmd[5:5] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)              # atomic: a process that already mapped the old library keeps it
    with open(stamp, "w") as o:
        o.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
