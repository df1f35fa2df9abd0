"""Builds libpvsr.so (sm_100a) in-tree with nvcc. Usage: python build.py [--force]

The shared library is the drop-in boundary (include/pvsr.h). It is built with an explicit
-gencode for sm_100a (tcgen05/TMA need the arch-specific target) and a static CUDA runtime.
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["conv3x3_tc.cu", "wgrad.cu", "simt_kernels.cu", "simt_bwd.cu", "tail_rank1.cu", "data_kernels.cu", "drf_kernels.cu", "tensormap.cpp", "api.cpp", "plan.cpp"]
HEADERS = ["conv.h", "wgrad.h", "ptx.cuh", "stencil.cuh", "simt.h", "internal.h", os.path.join("..", "..", "include", "pvsr.h")]
OUT = os.path.join(HERE, "libpvsr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-cudart", "static"]


STAMP = os.path.join(HERE, "libpvsr.build.json")   # travels with the .so (git-ignored): what the binary was built from
LAST = {"compiled": 0, "reused": None, "source_hash": None}   # record of the last build() call in this process


def source_hash():
    """sha256 over the contents of every source / header and the nvcc flags: the identity of a build."""
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        path = os.path.join(HERE, f)
        if os.path.exists(path):
            h.update(f.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """True unless libpvsr.so exists AND its stamp names the current source hash (content-based, not mtime-based: a
    snapshot copied to another box keeps working, an edited source always rebuilds)."""
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    try:
        with open(STAMP) as f:
            return json.load(f).get("source_hash") != source_hash()
    except Exception:
        return True


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    sh = source_hash()
    if not force and not needs_build():
        LAST.update(compiled=0, reused=True, source_hash=sh)
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
    with open(STAMP, "w") as f:
        json.dump({"source_hash": sh, "sources": srcs, "flags": FLAGS, "nvcc": NVCC}, f, indent=1)
    LAST.update(compiled=len(srcs), reused=False, source_hash=sh)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(json.dumps(LAST))
