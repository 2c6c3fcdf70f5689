"""Builds libracb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU):

    python -m robot_aware_control_b200.build [--force] [-v]

Every .cu is compiled to its own object in parallel (one nvcc per source, objects cached under lib/obj by mtime of the
source and of every header), then linked into one shared library. No relocatable device code: the sources do not call
each other's device functions.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libracb200.so")
OBJ = os.path.join(HERE, "lib", "obj")
SOURCES = ["conv_tc.cu", "conv_tc2.cu", "conv_tc_mc.cu", "conv_halo.cu", "first_conv_tc.cu", "misc_kernels.cu",
           "cem_kernels.cu", "train_kernels.cu", "train_gn_kernels.cu", "norm_lstm.cu", "metric_kernels.cu",
           "data_kernels.cu", "robot_kernels.cu", "wgrad_tc.cu", "rac_api.cu"]
# every header of csrc/ (a struct that crosses object files, e.g. WgradGeom, must rebuild all of its users)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + ["rac_train.inc.cu",
                                                                       os.path.join("..", "..", "include", "racb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS)


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in _sources() + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    hdr_t = _newest_header()
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(CSRC, src)), hdr_t)):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        logs[src] = res.stderr
        return obj

    with ThreadPoolExecutor(max_workers=min(len(_sources()), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        for s in _sources():
            if s in logs:
                print(f"==== {s}\n{logs[s]}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
