"""TEST INFRASTRUCTURE: build oracle/_ref/libdcn_ref.so from the reference's own dcn kernel source.

    python oracle/build_dcn_ref.py            # needs /root/reference (build container only)

The reference's extension (dcn/setup.py -> deform_conv_cuda) cannot be built against torch 2.11 (46 uses of removed
THC / AT_CHECK / .data<T>() host APIs in deform_conv_cuda.cpp).  Its KERNEL file, however, uses only three ATen names;
with oracle/dcn_ref_shim/ on the include path it compiles unmodified with plain nvcc.  The .so is git-ignored and
travels to the GPU box with gpurun; without /root/reference (on the box) this script is a no-op and the tests that
need the library skip when it is absent.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_KERNEL = "/root/reference/dcn/src/deform_conv_cuda_kernel.cu"
OUT = os.path.join(HERE, "_ref", "libdcn_ref.so")


def build(force=False):
    if not os.path.exists(REF_KERNEL):
        return OUT if os.path.exists(OUT) else None
    srcs = [REF_KERNEL, os.path.join(HERE, "dcn_ref_wrap.cu"), os.path.join(HERE, "dcn_ref_shim", "ATen", "ATen.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(HERE, "dcn_ref_shim"),
           "-DREF_KERNEL_FILE=\"%s\"" % REF_KERNEL, "-o", OUT, os.path.join(HERE, "dcn_ref_wrap.cu")]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
