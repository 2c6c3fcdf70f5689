"""Builds libracb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU):

    python -m robot_aware_control_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libracb200.so")
SOURCES = ["conv_tc.cu", "conv_tc2.cu", "conv_tc_mc.cu", "conv_halo.cu", "first_conv_tc.cu", "misc_kernels.cu", "cem_kernels.cu", "train_kernels.cu", "train_gn_kernels.cu", "norm_lstm.cu", "metric_kernels.cu", "data_kernels.cu", "rac_api.cu"]
HEADERS = ["conv.cuh", "epilogue.cuh", "ptx.cuh", "misc_kernels.cuh", "train_kernels.cuh", "rac_train.inc.cu", os.path.join("..", "..", "include", "racb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
