"""Build librvh.so (the C-ABI library, include/rvh.h) for sm_100a with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rvh_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "rvh_kernels.cuh"), os.path.join(HERE, "csrc", "rvh_host_math.h"),
        os.path.join(os.path.dirname(HERE), "include", "rvh.h")]
OUT = os.path.join(HERE, "librvh.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CCBIN = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


HOST_SRC = [os.path.join(HERE, "host", "rvh_host.cpp"), os.path.join(HERE, "host", "rvh_host_capi.cpp")]
HOST_DEPS = HOST_SRC + [os.path.join(HERE, "host", "rvh_host.hpp"), os.path.join(os.path.dirname(HERE), "include", "rvh.h")]
HOST_OUT = os.path.join(HERE, "librvh_host.so")


def build_host(force=False):
    """librvh_host.so: the C++ Hair/Scene/Renderer mirror (host/), linked against librvh.so next to it."""
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_OUT) for d in HOST_DEPS + [OUT]):
        return HOST_OUT
    cmd = [CCBIN, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", HOST_OUT] + HOST_SRC + \
          ["-L" + HERE, "-lrvh", "-Wl,-rpath,$ORIGIN"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return HOST_OUT


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_host()
        return OUT
    cmd = [NVCC, "-ccbin", CCBIN, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3", "-o", OUT, SRC, "-ldl"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    build_host(force=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
