"""Build libiblnerf_b200.so in-tree with nvcc for sm_100a (no torch headers involved: the boundary is a C ABI)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libiblnerf_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "iblnerf_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


LIB_DIAG = os.path.join(HERE, "libiblnerf_b200_diag.so")


def build_library(force=False, verbose=False, diag=False):
    """diag=True builds the tuning library libiblnerf_b200_diag.so (same sources + -DIBLN_DIAGNOSTICS: bandwidth
    probes, clock64 timelines, launch-skip switches; include/iblnerf_b200_diag.h) -- never loaded by the product."""
    if diag:
        return _build(LIB_DIAG, os.path.join(HERE, "build_diag"), ["-DIBLN_DIAGNOSTICS"], verbose)
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), [], verbose)


def _build(LIB, odir, extra, verbose):
    objs = []
    procs = []
    os.makedirs(odir, exist_ok=True)
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True, diag="--diag" in sys.argv))
