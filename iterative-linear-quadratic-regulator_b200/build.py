"""Build libilqr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "ilqr_b200.cu")]
HEADERS = ["ilqr_systems.cuh", "ilqr_trig_table.cuh", "ilqr_kernels_common.cuh", "ilqr_kernels_linearize.cuh", "ilqr_kernels_backward.cuh", "ilqr_kernels_ltv_mma.cuh",
           "ilqr_kernels_fused.cuh", "ilqr_kernels_rollout.cuh"]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in HEADERS] + [os.path.join(ROOT, "include", "ilqr_b200.h")]
OUT = os.path.join(HERE, "libilqr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
