"""CPU: the C-ABI shared library builds, loads and exports every symbol include/lsim.h declares; descriptor structs
in the ctypes binding match the header's layout.  No compute calls (no GPU here)."""

import ctypes as C
import os
import re
import subprocess

import pytest

from jaxtronomy_b200 import _lib
from jaxtronomy_b200.csrc import build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lsim.h")


@pytest.fixture(scope="module")
def lib():
    build_mod.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsim_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "liblsim.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "binding list and header disagree"


def test_abi_version_and_strerror(lib):
    assert lib.lsim_abi_version() == _lib.ABI_VERSION
    assert lib.lsim_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.lsim_strerror(code)) > 3
    assert lib.lsim_launch_count() == 0


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof as the C compiler sees them == what ctypes computes."""
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "lsim.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\\n", sizeof(lsim_lens_desc), offsetof(lsim_lens_desc, cst), offsetof(lsim_lens_desc, opt),
         offsetof(lsim_lens_desc, delta_T), offsetof(lsim_lens_desc, red2phys));
  printf("%zu %zu %zu\\n", sizeof(lsim_light_desc), offsetof(lsim_light_desc, col), offsetof(lsim_light_desc, cst));
  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(lsim_plan_desc), offsetof(lsim_plan_desc, lens),
         offsetof(lsim_plan_desc, light), offsetof(lsim_plan_desc, x_sub), offsetof(lsim_plan_desc, data),
         offsetof(lsim_plan_desc, mask));
  printf("%zu %zu %zu\\n", offsetof(lsim_plan_desc, eval_mask), offsetof(lsim_plan_desc, kernel_hi),
         offsetof(lsim_plan_desc, logl_offset));
  printf("%zu %zu %zu %zu\\n", offsetof(lsim_plan_desc, precision), offsetof(lsim_plan_desc, points),
         offsetof(lsim_plan_desc, ps_Ma2pix), offsetof(lsim_plan_desc, ps_scale));
  printf("%zu %zu %zu\\n", sizeof(lsim_point_desc), offsetof(lsim_point_desc, cst_ra), offsetof(lsim_point_desc, use_amp));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = [int(v) for v in out]
    L, G, P, Q = _lib.LensDesc, _lib.LightDesc, _lib.PlanDesc, _lib.PointDesc
    want = [C.sizeof(L), L.cst.offset, L.opt.offset, L.delta_T.offset, L.red2phys.offset,
            C.sizeof(G), G.col.offset, G.cst.offset,
            C.sizeof(P), P.lens.offset, P.light.offset, P.x_sub.offset, P.data.offset, P.mask.offset,
            P.eval_mask.offset, P.kernel_hi.offset, P.logl_offset.offset,
            P.precision.offset, P.points.offset, P.ps_Ma2pix.offset, P.ps_scale.offset,
            C.sizeof(Q), Q.cst_ra.offset, Q.use_amp.offset]
    assert got == want


def test_null_arguments_are_rejected_without_a_gpu(lib):
    handle = C.c_void_p()
    assert lib.lsim_plan_create(None, C.byref(handle)) == -1
    d = _lib.PlanDesc()
    d.abi_version = 999
    assert lib.lsim_plan_create(C.byref(d), C.byref(handle)) == -1
    d.abi_version = _lib.ABI_VERSION
    d.n_slots = 0
    assert lib.lsim_plan_create(C.byref(d), C.byref(handle)) == -1
    assert lib.lsim_workspace_bytes(None, 10) == 0
    assert lib.lsim_ray_shoot(None, None, 1, None, None, 1, 0, None, None, 0, None) == -1
    assert lib.lsim_logl(None, None, 1, 0, 0, None, None, None, None, 0, None) == -1
    assert lib.lsim_hessian(None, None, 1, None, None, 1, 0, None, None) == -1
    assert lib.lsim_source_position_logl(None, None, 1, None, None, 4, 0, 0.001, -1.0, None, None) == -1


def test_product_has_no_cpu_fallback():
    """Without a CUDA device every compute entry point must raise, never silently compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    lm = LensModel(["SIS"])
    with pytest.raises(_lib.LsimError):
        lm.ray_shooting([0.5], [0.2], [dict(theta_E=1.0, center_x=0.0, center_y=0.0)])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jaxtronomy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_xla_ffi_shim_is_header_gated(tmp_path):
    """csrc/lsim_xla_ffi.cc must compile without the jax headers (to the one-symbol stub) so that it can sit in
    the tree of a box that has no jax; with the headers it defines the XLA-FFI handlers."""
    import subprocess

    src = os.path.join(ROOT, "jaxtronomy_b200", "csrc", "lsim_xla_ffi.cc")
    so = str(tmp_path / "shim.so")
    subprocess.check_call(["g++", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), src, "-o", so])
    h = C.CDLL(so)
    assert h.lsim_xla_ffi_unavailable() == 1
    text = open(src).read()
    for handler in ("LsimLogl", "LsimLoglGrad", "LsimImage", "LsimRayShoot", "LsimHessian"):
        assert "XLA_FFI_DEFINE_HANDLER_SYMBOL(%s" % handler in text
