"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200pmp.h
declares, shares its structs with the oracle byte-for-byte, validates arguments, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import O, ROOT, _capi, as_oracle_opts, c_problem, oracle_problem
from approximate_optimal_control_b200 import systems


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200pmp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    declared = header_functions()
    assert set(declared) == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.pmp_abi_version() == _capi.ABI_VERSION


def test_header_struct_layout_matches_ctypes():
    """sizeof / field order of the ctypes mirrors equal what a C compiler lays out for the header."""
    import subprocess, tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200pmp.h"
int main(void){
  printf("%zu %zu %zu\n", sizeof(pmp_problem_t), sizeof(pmp_opts_t), sizeof(pmp_dense_t));
  printf("%zu %zu %zu %zu\n", offsetof(pmp_problem_t,u_lo), offsetof(pmp_problem_t,params), offsetof(pmp_opts_t,t0), offsetof(pmp_opts_t,event_level));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(f"{d}/a.c", "w").write(prog)
        subprocess.run(["/usr/bin/gcc", "-I", f"{ROOT}/include", f"{d}/a.c", "-o", f"{d}/a"], check=True)
        out = subprocess.run([f"{d}/a"], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:3] == [C.sizeof(_capi.Problem), C.sizeof(_capi.Opts), C.sizeof(_capi.Dense)]
    assert sizes[3:] == [_capi.Problem.u_lo.offset, _capi.Problem.params.offset, _capi.Opts.t0.offset,
                         _capi.Opts.event_level.offset]
    assert C.sizeof(O.Problem) == C.sizeof(_capi.Problem) and C.sizeof(O.Opts) == C.sizeof(_capi.Opts)


@pytest.mark.parametrize("name", ["flatquad", "orbits", "lqr"])
def test_problem_records_identical_for_oracle_and_product(name):
    prob, _ = c_problem(name)
    assert bytes(prob) == bytes(oracle_problem(name))
    o = _capi.make_opts()
    assert bytes(as_oracle_opts(o)) == bytes(o) == bytes(O.make_opts())


def test_state_rows_and_workspace():
    L = _capi.lib()
    prob, _ = c_problem("flatquad")
    assert L.pmp_state_rows(C.byref(prob)) == 14
    o = _capi.make_opts(dtype=_capi.F32)
    assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 1000) == 256 + 13 * 1000 * 4
    o = _capi.make_opts(dtype=_capi.F64)
    assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 1000) == 256 + 13 * 1000 * 8
    assert L.pmp_solve_launch_count(C.byref(prob), C.byref(o), 10) == 2
    assert L.pmp_solve_launch_count(C.byref(prob), C.byref(o), 0) == 0


def test_vxx_flag_rows_workspace_and_validation():
    """PMP_FLAG_VXX: 36 more rows for flatquad, refused for the nx = 2 functors, for value-parameterised solves
    and together with the literal u* form; the vxx event and the symmetric storage need the flag."""
    L = _capi.lib()
    prob, _ = c_problem("flatquad")
    V, S, LIT = _capi.FLAG_VXX, _capi.FLAG_VXX_SYMMETRIC, _capi.FLAG_USTAR_LITERAL
    assert L.pmp_state_rows_flags(C.byref(prob), 0) == 14 and L.pmp_state_rows_flags(C.byref(prob), V) == 50
    for flags in (V, V | S):
        o = _capi.make_opts(dtype=_capi.F64, flags=flags)
        assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 100) == 256 + (13 + 36) * 100 * 8
    orb, _ = c_problem("orbits")
    assert L.pmp_state_rows_flags(C.byref(orb), V) == -2
    assert L.pmp_solve_workspace_bytes(C.byref(orb), C.byref(_capi.make_opts(flags=V)), 8) == -2
    assert b"flatquad" in L.pmp_last_error()
    for kw, frag in ((dict(flags=S), b"needs PMP_FLAG_VXX"), (dict(flags=V | LIT), b"KKT"),
                     (dict(flags=V, indep=_capi.INDEP_VALUE, dt0=0.1), b"time-parameterised"),
                     (dict(event=_capi.EVENT_VXX_NORM), b"needs PMP_FLAG_VXX"), (dict(flags=8), b"unknown flags")):
        o = _capi.make_opts(**kw)
        assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 8) < 0, kw
        assert frag in L.pmp_last_error(), (kw, L.pmp_last_error())
    # dense output with the flag needs the sixth leaf
    o = _capi.make_opts(flags=V, save_dense=1)
    buf = (C.c_float * (50 * 8))()
    d = _capi.Dense(*(C.addressof(buf),) * 5, None)
    assert L.pmp_solve_batch_cuda(C.byref(prob), C.byref(o), 8, buf, buf, C.byref(d), None, None, None, None, 0, None) == -1
    assert b"vxx leaf" in L.pmp_last_error()


def test_argument_validation_and_error_strings():
    L = _capi.lib()
    prob, _ = c_problem("flatquad")
    bad = _capi.Problem(system_id=7, nx=6, nu=2)
    assert L.pmp_state_rows(C.byref(bad)) == -2 and b"unknown system_id" in L.pmp_last_error()
    bad = _capi.Problem.from_buffer_copy(bytes(prob)); bad.nx = 4
    assert L.pmp_state_rows(C.byref(bad)) == -1
    bad = _capi.Problem.from_buffer_copy(bytes(prob)); bad.u_hi[0] = -1.0
    assert L.pmp_state_rows(C.byref(bad)) == -1 and b"U_interval" in L.pmp_last_error()
    for kw, frag in ((dict(method=9), b"method"), (dict(max_steps=0), b"max_steps"), (dict(dt0=0.0), b"dt0"),
                     (dict(dt0=+0.1), b"dt0"), (dict(method="rk4", adaptive=1), b"RK4"), (dict(rtol=-1.0), b"rtol"),
                     (dict(dtype=5), b"dtype"), (dict(event=4), b"event"), (dict(event=3), b"pmp_forward_batch_cuda")):
        o = _capi.make_opts(**kw)
        assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 8) < 0, kw
        assert frag in L.pmp_last_error(), (kw, L.pmp_last_error())
    o = _capi.make_opts()
    # NULL buffers / missing workspace are rejected before anything touches a device
    assert L.pmp_solve_batch_cuda(C.byref(prob), C.byref(o), 8, None, None, None, None, None, None, None, 0, None) == -1
    buf = (C.c_float * (14 * 8))()
    assert L.pmp_solve_batch_cuda(C.byref(prob), C.byref(o), 8, buf, buf, None, None, None, None, None, 0, None) == -1
    assert b"workspace" in L.pmp_last_error()
    assert L.pmp_solve_batch_cuda(C.byref(prob), C.byref(o), 0, None, None, None, None, None, None, None, 0, None) == 0
    with pytest.raises(_capi.PmpError):
        _capi.check(-1)


def test_solve_host_validation_and_plan():
    """pmp_solve_host: argument checks run before any CUDA call; the plan (pieces, kernels, copies) is host logic."""
    L = _capi.lib()
    prob, _ = c_problem("flatquad")
    o = _capi.make_opts()
    nul = [None] * 11
    assert L.pmp_solve_host(C.byref(prob), C.byref(o), -1, *nul, 0) == -1
    assert L.pmp_solve_host(C.byref(prob), C.byref(o), 0, *nul, 0) == 0
    assert L.pmp_solve_host(C.byref(prob), C.byref(o), 8, *nul, 0) == -1 and b"x_f" in L.pmp_last_error()
    assert L.pmp_solve_host(C.byref(prob), C.byref(o), 8, *nul, -2) == -1 and b"n_pieces" in L.pmp_last_error()
    for kw in (dict(save_dense=1), dict(flags=_capi.FLAG_VXX)):
        ob = _capi.make_opts(**kw)
        assert L.pmp_solve_host(C.byref(prob), C.byref(ob), 8, *nul, 0) == -2 and b"endpoints only" in L.pmp_last_error()
    k, c = C.c_int32(), C.c_int32()
    for n, req, pieces in ((1 << 20, 0, 11), (1 << 18, 0, 5), (1 << 17, 0, 2), (1000, 0, 1), (1 << 20, 3, 3), (100, 8, 1), (0, 0, 0)):
        assert L.pmp_solve_host_plan(C.byref(prob), C.byref(o), n, req, C.byref(k), C.byref(c)) == pieces, (n, req)
        persistent = n >= (1 << 16)  # one persistent launch of the integrator instead of pack + prep + solve + unpack per piece
        assert k.value == (1 if persistent else 4 * pieces) and c.value == (10 if persistent else 9) * pieces, (n, k.value, c.value)
    big = _capi.make_opts(max_steps=1 << 14)  # counters no longer fit the packed word: three copies instead of one
    assert L.pmp_solve_host_plan(C.byref(prob), C.byref(big), 1 << 20, 4, C.byref(k), C.byref(c)) == 4 and c.value == 48
    assert L.pmp_host_release() == 0  # nothing allocated: a no-op


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _capi.lib()
    assert L.pmp_device_count() == 0
    prob, pp = c_problem("flatquad")
    y = np.zeros((14, 8), np.float32)
    assert L.pmp_rhs_batch_cuda(C.byref(prob), _capi.F32, 8, y.ctypes.data, y.ctypes.data, None) == _capi.lib().pmp_rhs_batch_cuda(
        C.byref(prob), _capi.F32, 8, y.ctypes.data, y.ctypes.data, None) == -4
    assert b"no CPU fallback" in L.pmp_last_error()
    x = np.zeros((8, 6), np.float32)
    o = _capi.make_opts()
    assert L.pmp_solve_host(C.byref(prob), C.byref(o), 8, x.ctypes.data, None, x.ctypes.data, x.ctypes.data, *([None] * 7), 0) == -4
    from approximate_optimal_control_b200 import pontryagin_utils as pu
    solve_backward, f_extended = pu.define_backward_solver(pp, systems.flatquad_algo_params())
    st = {"x": np.zeros(6, np.float32), "t": 0, "v": np.float32(0), "vx": np.zeros(6, np.float32)}
    with pytest.raises(_capi.PmpError):
        solve_backward(st)
    with pytest.raises(_capi.PmpError):
        pu.u_star_2d(st["x"], st["vx"], pp)


def test_product_does_not_touch_the_oracle():
    """Nothing under the package may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "approximate_optimal_control_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt.replace("oracle/opcount.hpp; u*", "") and "import oracle" not in txt \
                    and "from oracle" not in txt and "pmp_oracle" not in txt, os.path.join(dirpath, f)
    import subprocess
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libb200pmp.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
