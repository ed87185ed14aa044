"""Build libnbp_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build())."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libnbp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH
# geometry kernels are pinned op-for-op to the fp32 oracle: forbid FMA contraction there
SOURCES = {
    "abi.cu": [],
    "raster.cu": ["-fmad=false"],
    "backproject.cu": ["-fmad=false"],
    "scatter.cu": ["-fmad=false"],
    "planner.cu": ["-fmad=false"],
    "coverage.cu": ["-fmad=false"],
    "collision.cu": ["-fmad=false"],
    "section.cu": ["-fmad=false"],
    "conv_tc.cu": [],
    "nn_kernels.cu": [],
    "train_kernels.cu": [],
    "wgrad_tc.cu": [],
}


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(verbose: bool = False, force: bool = False) -> str:
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "nbp_b200.h"))
    objs, rebuilt = [], False
    procs = []
    for src, extra in SOURCES.items():
        s = os.path.join(HERE, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [NVCC] + COMMON + extra + ["-Xptxas", "-v", "-c", s, "-o", o]
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, " ".join(cmd), out))
        if verbose:
            print("==", src)
            print(out)
    if rebuilt or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
