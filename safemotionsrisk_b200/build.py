"""Builds libsmenv.so in-tree with nvcc for sm_100a (the GPU box uses the prebuilt file that travels with the repo)."""
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.environ.get("SMENV_LIB") or os.path.join(CSRC, "libsmenv.so")  # SMENV_LIB: experiment builds
SOURCES = ["smenv.cu", "smenv_device.cuh", "smenv_geom.cuh", "smenv_kernels.cuh", "smenv_joint.cuh", "smenv_step.cuh",
           "smenv_gjk.cuh", "smenv_plan.cuh", "smenv_mlp.cuh",
           "smenv_pools.cuh", "smenv_human.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    hdr = os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "smenv.h")
    newest = max(os.path.getmtime(p) for p in [hdr] + [os.path.join(CSRC, s) for s in SOURCES])
    return os.path.getmtime(LIB) < newest


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("SMENV_NVCC_EXTRA", "").split()   # experiment builds: e.g. -DSM_CONTACT_MIN_BLOCKS=3 with SMENV_LIB set
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "smenv.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
