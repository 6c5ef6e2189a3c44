"""Build libcfen_b200.so in-tree with nvcc for sm_100a (no torch headers, no cuBLAS/cuDNN).

    python cfen_vit_dehazing_b200/csrc/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cfen_b200.cu")
OUT = os.path.join(HERE, "libcfen_b200.so")
DEPS = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))] + [
    os.path.join(HERE, "..", "..", "include", "cfen_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False, timeline=False):
    """timeline=True builds the instrumented variant libcfen_b200_tl.so (-DCFEN_TIMELINE: device-side clock
    stamps in the fused token kernels, tools/timeline_tail16.py); never loaded by the product path."""
    out = OUT.replace(".so", "_tl.so") if timeline else OUT
    if not force and not timeline and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-DCFEN_NO_FAST_MATH", "-diag-suppress", "179,39", "-Xcompiler", "-fPIC", "-shared", "-o", out, SRC]
    if timeline:
        cmd.insert(1, "-DCFEN_TIMELINE")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timeline="--timeline" in sys.argv)
