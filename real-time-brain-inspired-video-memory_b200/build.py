"""Builds libvidmem.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python real-time-brain-inspired-video-memory_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libvidmem.so")
SOURCES = ["api.cu", "store.cu", "select.cu", "scan_simt.cu", "scan_tc.cu", "pairs.cu", "arena.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None, objdir: str = None) -> str:
    """extra_flags / out / objdir: A/B builds (e.g. -DVIDMEM_TRIAGE_KERNELS) into another .so, selected at run time
    with VIDMEM_LIB=<path>; the shipped library is always the default build."""
    global OBJ, LIB, NVCC_FLAGS
    saved = (OBJ, LIB, NVCC_FLAGS)
    if out:
        OBJ, LIB, NVCC_FLAGS = objdir or (out + ".obj"), out, NVCC_FLAGS + list(extra_flags)
    try:
        return _build(force, verbose)
    finally:
        OBJ, LIB, NVCC_FLAGS = saved


def _build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "vidmem.h"))
    nvcc = _nvcc()
    objs, rebuilt = [], False
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = op + ".sha"
        dig = _digest([sp] + headers)
        objs.append(op)
        if not force and os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", op]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        rebuilt = True
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(OBJ, src + ".log"), "w") as f:
            f.write(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            sys.stderr.write(out)
        with open(stamp, "w") as f:
            f.write(dig)
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, extra_flags=defs, out=outs[0] if outs else None))
