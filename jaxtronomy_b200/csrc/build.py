"""Build jaxtronomy_b200/csrc/liblsim.so for sm_100a with nvcc (cross-compiles without a GPU).

    python jaxtronomy_b200/csrc/build.py [--force] [--verbose]
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["lsim_api.cu"]
DEPS = ["lsim_api.cu", "lsim_kernels.cuh", "lsim_profiles.cuh", "lsim_math.cuh", "lsim_f32.cuh", "lsim_cse_tables.h",
        "lsim_math_tables.h", "lsim_grad.cuh", "lsim_grad_kernels.cuh",
        os.path.join(ROOT, "include", "lsim.h")]
OUT = os.path.join(HERE, "liblsim.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--split-compile", "8"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % OUT)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.check_call(cmd, cwd=HERE)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
