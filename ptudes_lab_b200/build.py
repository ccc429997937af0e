"""Build libptk.so (sm_100a) in-tree with nvcc.  `python -m ptudes_lab_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PTK_LIB_SUFFIX / PTK_NVCC_EXTRA: build and load a tuning variant next to the default library
# (e.g. PTK_LIB_SUFFIX=_kx3 PTK_NVCC_EXTRA="-DPTK_ICP_KX=3"; the experiment batches profiles/exp_r2*.sh use it)
LIB = os.path.join(CSRC, "libptk%s.so" % os.environ.get("PTK_LIB_SUFFIX", ""))
SOURCES = [os.path.join(CSRC, "ptk.cu"), os.path.join(CSRC, "ptk_ingest.cu"), os.path.join(CSRC, "ptk_ekf.cpp")]
HEADERS = [os.path.join(CSRC, "ptk_device.cuh"), os.path.join(CSRC, "ptk_canon.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "ptk.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact parity with the oracle: no fused multiply-add contraction, IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PTK_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libptk.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
