"""Builds liborlg.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG), "csrc")
LIB = os.path.join(PKG, "liborlg.so")
SOURCES = ["orlg_api.cu"]
HEADERS = ["orlg_kernels.cuh", "orlg_device.cuh", "orlg_deeprmsa_fast.cuh", "orlg_rollout.cuh", "orlg_policy.cuh", "orlg_step_wide.cuh", "orlg_wrappers.cuh", os.path.join("..", "..", "include", "orlg.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xptxas=-v", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "-cudart", "shared"]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    extra = os.environ.get("ORLG_NVCC_EXTRA", "").split()          # e.g. -DORLG_FAST_THREADS=64 (tuning experiments)
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building liborlg.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
