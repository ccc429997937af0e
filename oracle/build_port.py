"""Build oracle/libkiss_port.so from oracle/kiss_port.c with gcc (TEST INFRASTRUCTURE ONLY).

-ffp-contract=off: the canon has no fused multiply-add; -O3 without -ffast-math keeps IEEE
semantics; -fopenmp mirrors upstream's TBB parallel_for / parallel_reduce.

There is no oracle/_ref: /root/reference is pure Python calling the absent kiss-icp wheel, so
no file of the reference compiles into this path (DESIGN.md "Oracle").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kiss_port.c")
LIB = os.path.join(HERE, "libkiss_port.so")
# the timed variant: built ON the machine that runs bench.py (-march=native), never shipped
LIB_NATIVE = os.path.join(HERE, "libkiss_port_native.so")
BASE = ["-O3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas"]
# the library that travels with the repo must run on any x86-64 host of the last decade
FLAGS = BASE + ["-march=x86-64-v2"]
FLAGS_NATIVE = BASE + ["-march=native"]


def needs_build(lib=LIB):
    return (not os.path.exists(lib)) or os.path.getmtime(SRC) > os.path.getmtime(lib) or \
        os.path.getmtime(os.path.abspath(__file__)) > os.path.getmtime(lib)


def _compile(flags, lib):
    # the image exports CC=/opt/gcc/bin/gcc, a second gcc without libgomp: prefer the system one
    cands = [os.environ.get("PTK_ORACLE_CC"), "/usr/bin/gcc", "gcc", os.environ.get("CC")]
    cands = [c for c in cands if c]
    r = None
    for cc in cands:
        if os.path.sep in cc and not os.path.exists(cc):
            continue
        r = subprocess.run([cc] + flags + ["-o", lib, SRC, "-lm"], capture_output=True, text=True)
        if r.returncode == 0:
            break
    if r is None or r.returncode != 0:
        if r is not None:
            sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed building " + os.path.basename(lib))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return lib


def build(force=False):
    if not force and not needs_build(LIB):
        return LIB
    return _compile(FLAGS, LIB)


def build_native(force=False):
    """-O3 -march=native build for the CPU baseline legs of bench.py (`g++ -O3 -march=native`).
    Compiled on the host that times it; results stay bit-identical (no fast-math, no contraction)."""
    if not force and not needs_build(LIB_NATIVE):
        return LIB_NATIVE
    return _compile(FLAGS_NATIVE, LIB_NATIVE)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    if "--native" in sys.argv:
        print(build_native(force="--force" in sys.argv))
