"""Build the C-ABI shared library (radtxfr_b200/libradtxfr_b200.so) in-tree with nvcc for sm_100a.

    python -m radtxfr_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind extern "C" (include/radtxfr_b200.h); it does not link against torch.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libradtxfr_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# experiment hooks: RADTXFR_NVCC_DEFS="-DRTX_ROUNDS=1 ..." adds defines, RADTXFR_LIB overrides the output/load path
EXTRA_DEFS = os.environ.get("RADTXFR_NVCC_DEFS", "").split()
LIB = os.environ.get("RADTXFR_LIB", LIB)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(os.path.dirname(PKG), "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB)
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(PKG, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", *EXTRA_DEFS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", *ARCH, "-o", LIB + ".tmp", *objs]
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
