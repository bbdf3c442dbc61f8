"""Builds librae.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

``python -m relation_autoencoder_b200.build`` or ``build()`` from ``__graft_entry__``.  nvcc cross-compiles for sm_100a
without a GPU.  The built .so sits next to this file so it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librae.so")
SOURCES = ["rae_engine.cu", "rae_encoder.cu", "rae_decoder_simt.cu", "rae_decoder_tc.cu", "rae_update.cu", "rae_sort.cu", "rae_peer.cu", "rae_sampler.cu"]
HEADERS = ["rae_common.cuh", "rae_internal.h", os.path.join("..", "..", "include", "rae.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librae.so cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + FLAGS + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(OBJ, src + ".log"), "w") as f:
            f.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link of librae.so failed")
    return LIB


def build_trace() -> str:
    """librae_trace.so at the repository root: the same sources with -DRAE_TRACE (in-kernel timeline of the tcgen05
    kernels, profiles/trace_tc.py).  A measurement build: never loaded by the package."""
    nvcc = _nvcc()
    out = os.path.join(os.path.dirname(HERE), "librae_trace.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [nvcc] + ARCH + [f for f in FLAGS if f not in ("-Xptxas", "-v")] + ["-DRAE_TRACE", "-shared", "-o", out] + srcs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("trace build failed")
    return out


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
