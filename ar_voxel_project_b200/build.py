"""Build libvoxcarve.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

The library is the product's only compute path; nothing here falls back to the CPU.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libvoxcarve.so")
SOURCES = ["voxcarve.cu"]
DEPS = ["voxcarve.cu", "vc_kernels.cuh", "mc_tables.inc", os.path.join("..", "..", "include", "voxcarve.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true",  # IEEE f32 divide/sqrt, denormals kept (reference is x86 SSE)
    "-Xptxas", "-v",
    "-ldl",  # NCCL is dlopen-ed by vc_comm_init (no load-time dependency); nvtx3 is header-only and dlopen-s its injection library
]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: experiment variants (-D switches) built next to the product library, never loaded by default"""
    if out is None and not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    target = out or LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvoxcarve.so")
    with open(target[:-3] + ".ptxas.log" if out else os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
