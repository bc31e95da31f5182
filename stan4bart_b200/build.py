"""Builds stan4bart_b200/libstan4bart_b200.so (sm_100a) in-tree with nvcc.

    python -m stan4bart_b200.build [--force]

The shared library exports the C ABI declared in include/stan4bart_b200.h.  No torch types
and no JIT cache are involved: the .so travels with the repo snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libstan4bart_b200.so")
# the dbarts C-callable table over the GPU sampler (glue/gpubart_shim.cpp: host-only C++, links against the library above)
ROOT = os.path.dirname(HERE)
SHIM_SRC = os.path.join(ROOT, "glue", "gpubart_shim.cpp")
SHIM_LIB = os.path.join(HERE, "libgpubart_shim.so")
SHIM_INCLUDES = ["-I", os.path.join(ROOT, "include", "dbarts_shim"), "-I", os.path.join(ROOT, "glue")]
SOURCES = ["bart.cu", "glmm.cu", "nuts.cu", "sampler.cu", "shard.cu"]
# host side: the NUTS control and the GLMM's O((K+q)^2) expansion run on the CPU ~1000 times per sweep; AVX2 + FMA is safe on
# every host a B200 sits in
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-mavx2", "-Xcompiler", "-mfma", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "stan4bart_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _shim_deps():
    out = [SHIM_SRC, os.path.join(ROOT, "glue", "gpubart_shim.h"), os.path.join(ROOT, "include", "stan4bart_b200.h")]
    d = os.path.join(ROOT, "include", "dbarts_shim", "dbarts")
    return out + [os.path.join(d, f) for f in os.listdir(d)]


def build_shim(force=False):
    """g++ -shared glue/gpubart_shim.cpp -> libgpubart_shim.so next to (and linked against) libstan4bart_b200.so."""
    if not force and os.path.exists(SHIM_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SHIM_LIB) for d in _shim_deps() + [LIB]):
        return SHIM_LIB
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra"] + SHIM_INCLUDES + [SHIM_SRC, "-o", SHIM_LIB,
           "-L", HERE, "-lstan4bart_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for the dbarts shim:\n{r.stdout}\n{r.stderr}")
    return SHIM_LIB


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_shim()
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_shim(force=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
