"""Build oracle/libkiss_port.so from oracle/kiss_port.c with gcc (TEST INFRASTRUCTURE ONLY).

-ffp-contract=off: the canon has no fused multiply-add; -O2 without -ffast-math keeps IEEE
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
FLAGS = ["-O2", "-march=x86-64-v2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
         "-Wall", "-Wno-unknown-pragmas"]


def needs_build():
    return (not os.path.exists(LIB)) or os.path.getmtime(SRC) > os.path.getmtime(LIB) or \
        os.path.getmtime(os.path.abspath(__file__)) > os.path.getmtime(LIB)


def build(force=False):
    if not force and not needs_build():
        return LIB
    # the image exports CC=/opt/gcc/bin/gcc, a second gcc without libgomp: prefer the system one
    cands = [os.environ.get("PTK_ORACLE_CC"), "/usr/bin/gcc", "gcc", os.environ.get("CC")]
    cands = [c for c in cands if c]
    for cc in cands:
        if os.path.sep in cc and not os.path.exists(cc):
            continue
        r = subprocess.run([cc] + FLAGS + ["-o", LIB, SRC, "-lm"], capture_output=True, text=True)
        if r.returncode == 0:
            break
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed building libkiss_port.so")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
