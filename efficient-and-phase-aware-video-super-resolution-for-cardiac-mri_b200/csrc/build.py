"""Builds libpvsr.so (sm_100a) in-tree with nvcc. Usage: python build.py [--force]

The shared library is the drop-in boundary (include/pvsr.h). It is built with an explicit
-gencode for sm_100a (tcgen05/TMA need the arch-specific target) and a static CUDA runtime.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["conv3x3_tc.cu", "wgrad.cu", "simt_kernels.cu", "simt_bwd.cu", "data_kernels.cu", "tensormap.cpp", "api.cpp", "plan.cpp"]
HEADERS = ["conv.h", "wgrad.h", "ptx.cuh", "stencil.cuh", "simt.h", "internal.h", os.path.join("..", "..", "include", "pvsr.h")]
OUT = os.path.join(HERE, "libpvsr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-cudart", "static"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", s.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(HERE, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
