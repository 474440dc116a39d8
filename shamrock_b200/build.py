"""Build libshamb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m shamrock_b200.build [--force]

Everything is compiled with -fmad=false (bit-exact integer and comparison results, float64 outputs
bit-identical to the CPU oracle) except sph2_fast.cu, the fast-fp variant of the SPH loops selected at
run time by shamb200_solver_config.fp_mode (FMA contraction on, 1e-10 relative contract).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libshamb200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]
SOURCES = {
    "runtime.cu": ["-fmad=false"],
    "microbench.cu": ["-fmad=true"],
    "tree.cu": ["-fmad=false"],
    "neigh.cu": ["-fmad=false"],
    "stream_kernels.cu": ["-fmad=false"],
    "sph.cu": ["-fmad=false"],
    "neigh2.cu": ["-fmad=false"],
    "sph2.cu": ["-fmad=false"],
    "sph2_strict.cu": ["-fmad=false"],
    "sph2_fast.cu": ["-fmad=true"],
    "solver.cu": ["-fmad=false"],
    "setup.cu": ["-fmad=false"],
    "scheduler.cu": ["-fmad=false"],
    "dump.cu": ["-fmad=false"],
    "io_formats.cu": ["-fmad=false"],
    "solver_comm.cu": ["-fmad=false"],
    "capi.cu": ["-fmad=false"],
    "capi_stages.cu": ["-fmad=false"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _host_cxx():
    # the image exports CXX=/opt/gcc/bin/g++ which lacks some runtime pieces; use the distro g++
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "shamb200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(item):
        src, flags = item
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, "-ccbin", _host_cxx()] + ARCH + COMMON + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES.items()))
    cmd = [nvcc, "-ccbin", _host_cxx()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
