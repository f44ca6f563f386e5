"""Plug-in dynamics (codegen.py), CPU side: the sympy derivation is pinned on the
reference-generated goldens, the reference's own f, l source traces unmodified through symnp, and the generated
plug-in library builds (nvcc cross-compiles without a GPU), loads and exports the whole C ABI."""
import ctypes as C
import os
import textwrap

import numpy as np
import pytest

from common import GOLDEN, ROOT, _capi
from approximate_optimal_control_b200 import codegen, symnp, systems
import codegen_systems as CS

REF = "/root/reference"


def test_derived_flatquad_expressions_match_reference_golden():
    """lambdified sympy derivatives of systems.flatquad f, l against values produced by the reference's own jax-style
    code (tests/golden/rhs_flatquad_f64.npz): f(x,u*), -l(x,u*), -dH/dx at the reference's u*; and the quadratic
    model (H_u, H_uu at u = 0) reproduces the reference's u* through an independent box QP."""
    g = np.load(f"{GOLDEN}/rhs_flatquad_f64.npz")
    pp = systems.flatquad_problem_params()
    d = codegen.derive(pp)
    model, rhs = codegen.lambdify_rhs(d)
    lo, hi = (np.asarray(b, dtype=float) for b in pp["U_interval"])
    for i in range(g["x"].shape[0]):
        f, l, hx = rhs(g["x"][i], g["u"][i], g["lam"][i])
        assert np.allclose(f, g["dx"][i], rtol=1e-12, atol=1e-12)
        assert np.isclose(-l, g["dv"][i], rtol=1e-12, atol=1e-12)
        assert np.allclose(-np.asarray(hx, dtype=float), g["dlam"][i], rtol=1e-12, atol=1e-11)
        Hu, Huu = model(g["x"][i], g["lam"][i])
        assert np.allclose(Huu, [0.325, -0.315, 0.325], rtol=1e-12)
        assert np.allclose(CS.box_qp_2d(Hu, Huu, lo, hi), g["u"][i], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("name,mk", [("orbits", systems.orbits_problem_params), ("lqr", systems.lqr_statepenalty_problem_params)])
def test_derived_nu1_expressions_match_reference_golden(name, mk):
    g = np.load(f"{GOLDEN}/rhs_{name}_f64.npz")
    pp = mk()
    d = codegen.derive(pp)
    model, rhs = codegen.lambdify_rhs(d)
    lo, hi = pp["U_interval"]
    for i in range(g["x"].shape[0]):
        num, den = model(g["x"][i], g["lam"][i])
        u = np.clip(num / den, lo, hi)
        assert np.allclose(u, g["u"][i], rtol=1e-13, atol=1e-14)  # u_star_new
        f, l, hx = rhs(g["x"][i], np.atleast_1d(g["u"][i]), g["lam"][i])
        assert np.allclose(f, g["dx"][i], rtol=1e-12, atol=1e-12) and np.isclose(-l, g["dv"][i], rtol=1e-12, atol=1e-12)
        assert np.allclose(-np.asarray(hx, dtype=float), g["dlam"][i], rtol=1e-12, atol=1e-11)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("path,lines,nx,nu,mk", [
    ("flatquad_landing_experiment.py", (269, 329), 6, 2, systems.flatquad_problem_params),
    ("experiment_orbits.py", (24, 40), 2, 1, systems.orbits_problem_params),
    ("experiment_lqr_statepenalty.py", (23, 38), 2, 1, systems.lqr_statepenalty_problem_params)])
def test_reference_source_traces_unmodified(path, lines, nx, nu, mk):
    """exec the reference's OWN f, l definitions (read from /root/reference at test time, nothing copied) with
    symnp standing in for jax.numpy: the derived expressions equal those of the package's transcription."""
    import sympy as sp
    src = "".join(open(os.path.join(REF, path)).readlines()[lines[0] - 1:lines[1]])
    ns = {"np": symnp}
    exec(compile(textwrap.dedent(src), path, "exec"), ns)
    a = codegen.derive({"f": ns["f"], "l": ns["l"], "nx": nx, "nu": nu})
    b = codegen.derive(mk())
    ea = a.f + [a.l] + a.dHdx + (a.Hu + a.Huu if nu == 2 else [a.num, a.den])
    eb = b.f + [b.l] + b.dHdx + (b.Hu + b.Huu if nu == 2 else [b.num, b.den])
    rng = np.random.Generator(np.random.PCG64(3))
    syms = a.x + a.u + a.lam
    for _ in range(5):
        sub = dict(zip(syms, rng.standard_normal(len(syms))))
        for p, q in zip(ea, eb):
            vp, vq = float(sp.sympify(p).subs(sub)), float(sp.sympify(q).subs(sub))
            assert abs(vp - vq) <= 1e-11 * (1 + abs(vq))


def test_codegen_rejects_what_the_reference_formulas_do_not_cover():
    bad = dict(CS.pendulum_problem_params())
    bad["l"] = lambda x, u: x[0] ** 2 + u[0] ** 4  # H not quadratic in u
    with pytest.raises(NotImplementedError, match="quadratic"):
        codegen.derive(bad)
    with pytest.raises(NotImplementedError, match="nu"):
        codegen.derive({"f": lambda x, u: x, "l": lambda x, u: x[0], "nx": 2, "nu": 3})
    with pytest.raises(NotImplementedError, match="traced"):
        codegen.derive({"f": lambda x, u: np.sin(x), "l": lambda x, u: x[0], "nx": 2, "nu": 1})  # numpy, not symnp
    with pytest.raises(NotImplementedError):
        systems.to_c_problem({"system_name": "nope", "nx": 2, "nu": 1, "U_interval": [-1, 1]})  # no f, l


def test_generated_header_and_plugin_library():
    """generate + nvcc a plug-in (pendulum: nx = 2, nu = 1, exp / tanh in l); it exports every symbol of the header,
    answers to PMP_SYS_GENERATED only, and has no CPU fallback either."""
    pp = CS.pendulum_problem_params()
    hdr = codegen.emit_header(codegen.derive(pp), "pendulum_test")
    assert "struct Generated" in hdr and "gtanh" in hdr and "gexp" in hdr and "u_star_new" in hdr
    prob = systems.to_c_problem(pp)
    assert prob.system_id == _capi.SYS_GENERATED and os.path.exists(prob.lib_path)
    assert prob.lib_path == systems.to_c_problem(pp).lib_path  # cached by content hash
    L = _capi.lib_of(prob)
    for name in _capi.EXPORTS:
        assert hasattr(L, name), name
    assert L.pmp_state_rows(C.byref(prob)) == 6
    flat = systems.to_c_problem(systems.flatquad_problem_params())
    assert L.pmp_state_rows(C.byref(flat)) == -2 and b"plug-in" in L.pmp_last_error()
    assert _capi.lib().pmp_state_rows(C.byref(prob)) == -2  # and the main library does not know the generated id
    # vxx leaf of a generated functor (forward mode, csrc/pmp_vxx_dual.cuh): 2 nx + 1 + nx^2 workspace rows
    o = _capi.make_opts(dtype=_capi.F64, flags=_capi.FLAG_VXX)
    assert L.pmp_solve_workspace_bytes(C.byref(prob), C.byref(o), 8) == 256 + (5 + 4) * 8 * 8
    assert L.pmp_state_rows_flags(C.byref(prob), _capi.FLAG_VXX) == 6 + 4
    orb = systems.to_c_problem(systems.orbits_problem_params())
    assert _capi.lib().pmp_state_rows_flags(C.byref(orb), _capi.FLAG_VXX) == -2  # u_star_new systems: no vxx branch in the reference
    import torch
    if not torch.cuda.is_available():
        y = np.zeros((6, 8))
        assert L.pmp_rhs_batch_cuda(C.byref(prob), _capi.F64, 8, y.ctypes.data, y.ctypes.data, None) == -4


def test_feedback_law_is_traced_into_the_functor():
    """eval_utils.closed_loop_eval_general (eval_utils.py:48-99): a caller-supplied u*(x) becomes the generated functor's u*
    (no costate, no clipping, no quadratic-H requirement); its constants are part of the plug-in's content hash."""
    pp = dict(CS.pendulum_problem_params(), ustar_fct=CS.pendulum_feedback_law())
    d = codegen.derive(pp)
    assert d.policy is not None and len(d.policy) == 1 and d.num is None and not d.Hu
    hdr = codegen.emit_header(d, "pendulum_with_law")
    assert "the caller's feedback law" in hdr and "u_star_new" not in hdr and "ustar2d_box" not in hdr
    assert "nan_min<R>(R(2.0)" in hdr and "gsin(x[0])" in hdr  # clip(-(3 x0 + 1.5 x1) + 0.2 sin x0, -2, 2)
    f, l, u = CS.numpy_closed_loop(CS.pendulum_problem_params(), CS.pendulum_feedback_law())
    assert u(np.array([2.0, 0.0])) == -2.0 and abs(u(np.array([0.1, 0.0])) - (-0.3 + 0.2 * np.sin(0.1))) < 1e-15
    other = codegen.emit_header(codegen.derive(dict(CS.pendulum_problem_params(), ustar_fct=CS.lqr_feedback_law(np.eye(2)))), "pendulum_with_law")
    assert other != hdr
    with pytest.raises(ValueError, match="components"):
        codegen.derive(dict(CS.pendulum_problem_params(), ustar_fct=lambda x: x))  # returns nx = 2 values for nu = 1


def test_bind_host_to_gpu_never_fails():
    """sharding.bind_host_to_gpu reports what it did and leaves the process alone when the platform has nothing to say
    (here: no GPU at all)."""
    import os
    from approximate_optimal_control_b200 import sharding
    before = os.sched_getaffinity(0)
    info = sharding.bind_host_to_gpu(0)
    assert isinstance(info, dict) and "bound" in info
    import torch
    if not torch.cuda.is_available():
        assert info["bound"] is False and "why" in info and os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
