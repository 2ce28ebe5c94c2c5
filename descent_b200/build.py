"""Builds libdescent_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no pip install).

    python -m descent_b200.build [--force]

Sources: descent_b200/csrc/*.cpp (host, g++ via nvcc) and *.cu (device, -gencode arch=compute_100a,code=sm_100a).
The result descent_b200/libdescent_cuda.so is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
BUILD = os.path.join(ROOT, "build")
LIB = os.path.join(ROOT, "libdescent_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_LIB = "/usr/local/cuda/lib64"

COMMON = ["-O2", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I", os.path.join(ROOT, "..", "include")]
CU_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]


def _sources():
    out = []
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cpp", ".cu")):
            out.append(os.path.join(CSRC, name))
    return out


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "..", "include")):
        for name in os.listdir(d):
            if name.endswith((".hpp", ".h", ".cuh", ".inc")):  # kernel templates are #included by codegen.cpp
                m = max(m, os.path.getmtime(os.path.join(d, name)))
    return m


def _compile(src, obj):
    if src.endswith(".cu"):
        cmd = [NVCC, "-c", src, "-o", obj] + COMMON + CU_FLAGS
    else:
        cmd = ["g++", "-c", src, "-o", obj, "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-Wno-comment",
               "-I", os.path.join(ROOT, "..", "include"), "-I", "/usr/local/cuda/include"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stderr


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdr = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for warn in pool.map(lambda j: _compile(*j), jobs):
                if verbose and warn:
                    sys.stderr.write(warn)
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-L", CUDA_LIB, "-lnvrtc", "-ldl", "-lz", "-Xlinker", "-rpath," + CUDA_LIB, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
