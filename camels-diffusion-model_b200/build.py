"""Build libcdm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is compiled ahead of time so that it travels with the repo snapshot
to the GPU box; nothing is JIT-compiled at run time.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcdm_b200.so")
LIB_PROBES = os.path.join(HERE, "libcdm_b200_probes.so")  # same sources with -DCDM_PROBES (tools/gpu_probe.py only)
STAMP = os.path.join(HERE, ".libcdm_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "cdm_b200.h")])
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    jobs = [(LIB, []), (LIB_PROBES, ["-DCDM_PROBES"])]
    procs = []
    for lib, extra in jobs:  # the two libraries compile side by side
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-t", "4", "-o", lib] + _sources()
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for lib, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}")
        if verbose:
            sys.stderr.write(err)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
