"""Build jaxtronomy_b200/csrc/liblsim.so for sm_100a with nvcc (cross-compiles without a GPU).

    python jaxtronomy_b200/csrc/build.py [--force] [--verbose]
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# two translation units, compiled in parallel: the headline kernels alone and unsplit (lsim_hot_list.h says why), the API
# with everything else under --split-compile
SOURCES = [("lsim_hot.cu", []), ("lsim_api.cu", ["--split-compile", "8"])]
DEPS = ["lsim_api.cu", "lsim_hot.cu", "lsim_hot_list.h", "lsim_kernels.cuh", "lsim_profiles.cuh", "lsim_math.cuh", "lsim_f32.cuh", "lsim_cse_tables.h",
        "lsim_math_tables.h", "lsim_grad.cuh", "lsim_grad_kernels.cuh",
        os.path.join(ROOT, "include", "lsim.h")]
OUT = os.path.join(HERE, "liblsim.so")
# -static-global-template-stub=false: the kernels of lsim_hot.cu are launched from lsim_api.cu through `extern template`
# declarations, which needs externally visible host stubs in whole-program (non-rdc) mode
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-static-global-template-stub=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


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
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src, extra in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, cwd=HERE)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", OUT] + objs, cwd=HERE)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
