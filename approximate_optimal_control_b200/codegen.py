"""Plug-in dynamics: from the user's f(x,u), l(x,u) to a compiled CUDA functor.

The reference accepts ANY `problem_params['f']`, `['l']` because jax autodiff produces what the PMP right hand
side needs (pontryagin_utils.py:35-38 H_u, H_uu at u = 0; :438-439 dH/dlambda, dH/dx; :541-550 R, G of
u_star_new).  Here sympy takes jax's place: the same Python callables are traced with symbols (they must be
written against a numpy-like namespace that accepts sympy objects -- `approximate_optimal_control_b200.symnp`
mirrors the jax.numpy functions the reference's scripts use), differentiated symbolically, common
subexpressions are eliminated, and the result is printed as a `pmp::Generated<R>` device functor with the
interface of the hand-written ones in csrc/pmp_systems.cuh.  That header is compiled with the SAME hand-written
kernels (integrator, dense write-out, interpolation, pointwise evaluation) into a plug-in library
`plugins/libb200pmp_gen_<hash>.so` that exports the C ABI of include/b200pmp.h for system_id PMP_SYS_GENERATED.

    problem = codegen.plugin_problem(problem_params)     # trace, generate, nvcc (cached by content hash), load
    # or simply: problem_params without a compiled 'system_name' (or with problem_params['codegen'] = True)
    #            handed to pontryagin_utils.define_backward_solver / u_star_2d / ...

u* follows the reference: nu == 2 -> u_star_2d on the quadratic model of H at u = 0 (csrc/pmp_ustar2d.cuh),
nu == 1 -> u_star_new: clip(-(lambda' df/du) / (d2l/du2), U).  H must be quadratic in u (the reference assumes
it, pontryagin_utils.py:30-34); this is checked symbolically.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _capi
from . import build as _build

__all__ = ["derive", "emit_header", "build_plugin", "plugin_problem", "Derived", "lambdify_rhs"]

_PKG = os.path.dirname(os.path.abspath(__file__))
PLUGIN_DIR = os.path.join(_PKG, "plugins")
PLUGIN_SOURCES = ["pmp_capi.cu", "pmp_host.cu", "pmp_eval.cu", "pmp_select.cu", "pmp_gather.cu", "pmp_interp.cu", "pmp_forward.cu", "pmp_forward_nn.cu", "pmp_forward_cluster.cu", "pmp_forward_tc.cu", "pmp_forward_umma.cu", "pmp_sens.cu", "pmp_sys_generated.cu", "pmp_sys_generated_vxx.cu"]
MAX_NX = 16


# ---- 1. trace + differentiate ------------------------------------------------------------------------------
@dataclass
class Derived:
    nx: int
    nu: int
    x: list
    u: list
    lam: list
    f: list            # f_i(x, u)
    l: object          # l(x, u)
    dHdx: list         # dH/dx_i (x, u, lam), u held fixed (pontryagin_utils.py:439)
    Hu: list = field(default_factory=list)    # nu == 2: dH/du at u = 0
    Huu: list = field(default_factory=list)   # nu == 2: [h00, h01, h11] at u = 0
    num: object = None  # nu == 1: -lambda' df/du at u = 0
    den: object = None  # nu == 1: d2l/du2 at u = 0  (= R + R', R = 1/2 d2l/du2, pontryagin_utils.py:541-556)
    policy: object = None  # problem_params['ustar_fct']: a feedback law u(x) that takes the place of u* (eval_utils.py:48-99)


def _call(fn, x, u):
    import inspect
    try:
        nargs = len(inspect.signature(fn).parameters)
    except (TypeError, ValueError):
        nargs = 2
    return fn(x, u) if nargs == 2 else fn(0.0, x, u)  # stale 3-arg f(t, x, u) of the orbits / LQR scripts


def derive(problem_params: dict) -> Derived:
    import sympy as sp
    nx, nu = int(problem_params["nx"]), int(problem_params["nu"])
    if nu not in (1, 2):
        raise NotImplementedError("codegen: u_star_2d needs nu == 2, u_star_new nu == 1 (as the reference)")
    if not 1 <= nx <= MAX_NX:
        raise NotImplementedError(f"codegen: nx must be in 1..{MAX_NX} (one thread keeps a trajectory in registers)")
    x = [sp.Symbol(f"x[{i}]", real=True) for i in range(nx)]
    u = [sp.Symbol(f"u[{i}]", real=True) for i in range(nu)]
    lam = [sp.Symbol(f"lam[{i}]", real=True) for i in range(nx)]
    from .symnp import SymArray
    xa, ua = SymArray(x), SymArray(u)
    try:
        fr = np.asarray(_call(problem_params["f"], xa, ua), dtype=object).reshape(-1)
        lr = np.asarray(_call(problem_params["l"], xa, ua), dtype=object).reshape(-1)
    except Exception as e:
        raise NotImplementedError(
            "codegen: problem_params['f'] / ['l'] could not be traced with sympy symbols; write them against a "
            f"namespace that accepts sympy objects (approximate_optimal_control_b200.symnp): {type(e).__name__}: {e}") from e
    if fr.size != nx or lr.size != 1:
        raise ValueError(f"codegen: f returned {fr.size} components (nx = {nx}), l returned {lr.size}")
    def scalar(e):
        while isinstance(e, np.ndarray):  # 0-d object arrays left over from x[..., i] style indexing
            e = e.item()
        return sp.sympify(e)

    f = [scalar(e) for e in fr]
    l = scalar(lr[0])
    H = l + sum(li * fi for li, fi in zip(lam, f))
    if problem_params.get("ustar_fct") is not None:
        # closed_loop_eval_general(problem_params, algo_params, ustar_fct, x0s): "ustar_fct should be a jax function, X -> U. no
        # time dependence" (eval_utils.py:50-51).  Traced like f and l; the functor's u* becomes this law (the costate is ignored)
        try:
            ur = np.asarray(problem_params["ustar_fct"](xa), dtype=object).reshape(-1)
        except Exception as e:
            raise NotImplementedError("codegen: ustar_fct could not be traced with sympy symbols (write it against "
                                      f"approximate_optimal_control_b200.symnp): {type(e).__name__}: {e}") from e
        if ur.size != nu:
            raise ValueError(f"codegen: ustar_fct returned {ur.size} components (nu = {nu})")
        d = Derived(nx, nu, x, u, lam, f, l, [sp.diff(H, xi) for xi in x])
        d.policy = [scalar(e) for e in ur]
        return d
    # the reference's u* formulas assume H quadratic in u (pontryagin_utils.py:30-34)
    for i in range(nu):
        for j in range(i, nu):
            for k in range(j, nu):
                if sp.simplify(sp.diff(H, u[i], u[j], u[k])) != 0:
                    raise NotImplementedError("codegen: H = l + lambda'f is not quadratic in u; u_star_2d / u_star_new do not apply")
    zero = {ui: 0 for ui in u}
    d = Derived(nx, nu, x, u, lam, f, l, [sp.diff(H, xi) for xi in x])
    if nu == 2:
        d.Hu = [sp.diff(H, ui).subs(zero) for ui in u]
        d.Huu = [sp.diff(H, u[0], u[0]).subs(zero), sp.diff(H, u[0], u[1]).subs(zero), sp.diff(H, u[1], u[1]).subs(zero)]
    else:
        d.num = -sum(li * sp.diff(fi, u[0]).subs(zero) for li, fi in zip(lam, f))
        d.den = sp.diff(l, u[0], u[0]).subs(zero)
        if sp.simplify(d.den) == 0:
            raise NotImplementedError("codegen: d2l/du2 = 0 at u = 0: u_star_new would divide by zero")
    return d


def lambdify_rhs(d: Derived):
    """numpy callables of the derived expressions (tests / independent CPU checks):
    ustar(x, lam) for nu == 1 or (Hu, Huu)(x, lam) for nu == 2, and rhs(x, u, lam) -> (f, l, dHdx)."""
    import sympy as sp
    mods = ["numpy"]
    rhs = sp.lambdify([d.x, d.u, d.lam], [d.f, d.l, d.dHdx], modules=mods)
    if d.nu == 2:
        model = sp.lambdify([d.x, d.lam], [d.Hu, d.Huu], modules=mods)
    else:
        model = sp.lambdify([d.x, d.lam], [d.num, d.den], modules=mods)
    return model, rhs


# ---- 2. print as CUDA --------------------------------------------------------------------------------------
def _printer():
    from sympy.printing.c import C99CodePrinter
    from sympy.printing.precedence import PRECEDENCE
    import sympy as sp

    class P(C99CodePrinter):
        def _print_Float(self, e):
            return f"R({float(e)!r})"

        def _print_Integer(self, e):
            return f"R({int(e)})"

        def _print_Rational(self, e):
            return f"R({float(e)!r})"

        def _print_NumberSymbol(self, e):
            return f"R({float(e)!r})"

        _print_Pi = _print_Exp1 = _print_NumberSymbol

        def _print_Pow(self, e):
            base, ex = e.as_base_exp()
            b = self.parenthesize(base, PRECEDENCE["Mul"])
            if ex.is_Integer and 2 <= int(ex) <= 4:
                return "(" + "*".join([b] * int(ex)) + ")"
            if ex.is_Integer and -4 <= int(ex) <= -1:
                den = "*".join([b] * (-int(ex)))
                return f"M::div(R(1), {den})"
            if ex == sp.Rational(1, 2):
                return f"M::sqrt({self._print(base)})"
            if ex == -sp.Rational(1, 2):
                return f"M::div(R(1), M::sqrt({self._print(base)}))"
            return f"gpow({self._print(base)}, {self._print(ex)})"

        def _fn(self, name, e):
            return f"{name}({', '.join(self._print(a) for a in e.args)})"

        def _print_sin(self, e): return self._fn("gsin", e)
        def _print_cos(self, e): return self._fn("gcos", e)
        def _print_tan(self, e): return self._fn("gtan", e)
        def _print_tanh(self, e): return self._fn("gtanh", e)
        def _print_exp(self, e): return self._fn("gexp", e)
        def _print_log(self, e): return self._fn("glog", e)
        def _print_atan2(self, e): return self._fn("gatan2", e)
        def _print_Abs(self, e): return self._fn("M::abs", e)

        def _fold(self, name, args):
            out = self._print(args[0])
            for a in args[1:]:
                out = f"{name}<R>({out}, {self._print(a)})"
            return out

        def _print_Max(self, e): return self._fold("nan_max", e.args)   # jnp.maximum propagates NaN
        def _print_Min(self, e): return self._fold("nan_min", e.args)

        def _print_Heaviside(self, e):
            return f"(({self._print(e.args[0])}) > R(0) ? R(1) : R(0))"

        def _print_sign(self, e):
            a = self._print(e.args[0])
            return f"(({a}) > R(0) ? R(1) : (({a}) < R(0) ? R(-1) : R(0)))"

        def _print_Function(self, e):
            raise NotImplementedError(f"codegen: no CUDA printer for the function {e.func.__name__}")

        def _print_Piecewise(self, e):
            raise NotImplementedError("codegen: Piecewise expressions are not supported (use maximum / minimum)")

    return P()


def _cse_block(exprs, names, prefix, indent="        "):
    """CUDA statements computing `names[i] = exprs[i]` with common subexpressions in const temporaries."""
    import sympy as sp
    pr = _printer()
    repl, red = sp.cse(list(exprs), symbols=sp.numbered_symbols(prefix), optimizations="basic")
    lines = [f"{indent}const R {pr.doprint(s)} = {pr.doprint(e)};" for s, e in repl]
    lines += [f"{indent}{n} = {pr.doprint(e)};" for n, e in zip(names, red)]
    return "\n".join(lines), len(repl)


def emit_header(d: Derived, name: str = "user") -> str:
    nx, nu = d.nx, d.nu
    nd = 2 * nx + 1
    est32 = 9 * nd + 50  # stage derivatives (7) + state + stage state, + addressing / controller
    minb32 = int(min(4, max(1, 65536 // (128 * min(est32, 255)))))
    minb64 = int(min(3, max(1, 65536 // (128 * min(2 * est32, 255)))))
    if d.policy is not None:
        body_u, n_u = _cse_block(d.policy, [f"u[{i}]" for i in range(nu)], "a")
        ustar = body_u + "  // the caller's feedback law u(x) (closed_loop_eval_general, eval_utils.py:48-99): no costate, no clipping"
    elif nu == 2:
        body_u, n_u = _cse_block(d.Hu + d.Huu, ["const R Hu0", "const R Hu1", "const R h00", "const R h01", "const R h11"], "a")
        ustar = f"""{body_u}
        ustar2d_box<R, MODE>(Hu0, Hu1, h00, h01, h11, c.lo[0], c.lo[1], c.hi[0], c.hi[1], u[0], u[1]);"""
    else:
        body_u, n_u = _cse_block([d.num, d.den], ["const R num", "const R den"], "a")
        ustar = f"""{body_u}
        u[0] = clipf(M::div(num, den), c.lo[0], c.hi[0]);  // u_star_new: clip(solve(R + R', -lambda'G), U)"""
    outs = [f"xd[{i}]" for i in range(nx)] + ["const R l_"] + [f"const R hx{i}" for i in range(nx)]
    body_r, n_r = _cse_block(d.f + [d.l] + d.dHdx, outs, "b")
    neg = "\n".join(f"        ld[{i}] = -hx{i};" for i in range(nx))
    return f"""// GENERATED by approximate_optimal_control_b200/codegen.py -- do not edit.
// system '{name}': nx = {nx}, nu = {nu}; {n_u} + {n_r} common subexpressions.
// sympy derivatives of the user's f(x,u), l(x,u) in place of jax autodiff (pontryagin_utils.py:35-38,438-439,541-550).
#pragma once
#define PMP_GEN_NX {nx}
#define PMP_GEN_MINB32 {minb32}
#define PMP_GEN_MINB64 {minb64}
namespace pmp {{
template <class R> struct Generated {{
    static constexpr int NX = {nx}, NU = {nu};
    using M = Math<R>;
    struct Consts {{
        R lo[NU], hi[NU];
    }};
    static __host__ Consts make(const pmp_problem_t& p) {{
        Consts c;
        for (int i = 0; i < NU; i++) {{ c.lo[i] = R(p.u_lo[i]); c.hi[i] = R(p.u_hi[i]); }}
        return c;
    }}
    static __device__ __forceinline__ R gsin(R a) {{ R s, c; M::sincos(a, &s, &c); return s; }}
    static __device__ __forceinline__ R gcos(R a) {{ R s, c; M::sincos(a, &s, &c); return c; }}
    static __device__ __forceinline__ R gtan(R a) {{ R s, c; M::sincos(a, &s, &c); return M::div(s, c); }}
    static __device__ __forceinline__ R gexp(R a) {{ return fx::exp(a); }}
    static __device__ __forceinline__ R glog(R a) {{ return fx::log(a); }}
    static __device__ __forceinline__ R gtanh(R a) {{ return fx::tanh(a); }}
    static __device__ __forceinline__ R gatan2(R a, R b) {{ return fx::atan2(a, b); }}
    static __device__ __forceinline__ R gpow(R a, R b) {{ return fx::pow(a, b); }}

    // u* = argmin_u H over the box.  MODE & 2: KKT selection (integrator default), else the literal two-pass form
    template <int MODE>
    static __device__ __forceinline__ void ustar_mode(const Consts& c, const R* x, const R* lam, R* u) {{
{ustar}
    }}
    static __device__ __forceinline__ void ustar(const Consts& c, const R* x, const R* lam, R* u) {{ ustar_mode<0>(c, x, lam, u); }}

    // forward-time f_extended: xd = f(x,u*), vd = -l(x,u*), ld = -dH/dx at fixed u*
    template <int MODE>
    static __device__ __forceinline__ void rhs(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld) {{
        R u[NU];
        ustar_mode<MODE>(c, x, lam, u);
{body_r}
        vd = -l_;
{neg}
    }}
}};
}}  // namespace pmp
"""


# ---- 3. compile + load ------------------------------------------------------------------------------------
def _content_hash(header: str) -> str:
    h = hashlib.sha256()
    h.update(header.encode())
    csrc = os.path.join(_PKG, "csrc")
    for fn in sorted(os.listdir(csrc)):
        if fn.endswith((".cu", ".cuh")) and not fn.startswith("pmp_sys_f") and not fn.startswith("pmp_sys_o") and not fn.startswith("pmp_sys_l"):
            h.update(open(os.path.join(csrc, fn), "rb").read())
    h.update(open(os.path.join(os.path.dirname(_PKG), "include", "b200pmp.h"), "rb").read())
    h.update(" ".join(_build.NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build_plugin(problem_params: dict, force: bool = False, verbose: bool = False) -> str:
    """Generate + compile the plug-in library of problem_params (cached by content hash).  Returns its path."""
    name = str(problem_params.get("system_name", "user"))
    header = emit_header(derive(problem_params), name)
    tag = _content_hash(header)
    lib_path = os.path.join(PLUGIN_DIR, f"libb200pmp_gen_{tag}.so")
    if os.path.exists(lib_path) and not force:
        return lib_path
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    # one builder per tag at a time: with one process per GPU, N ranks meet here for the same new system; the others wait
    # for the lock and find the finished library
    import fcntl
    lock = open(os.path.join(PLUGIN_DIR, f"{tag}.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if os.path.exists(lib_path) and not force:
            return lib_path
        return _build_plugin_locked(header, tag, lib_path, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_plugin_locked(header: str, tag: str, lib_path: str, verbose: bool) -> str:
    work = os.path.join(PLUGIN_DIR, tag)
    os.makedirs(work, exist_ok=True)
    hpath = os.path.join(work, "pmp_generated.cuh")
    with open(hpath, "w") as fh:
        fh.write(header)
    nvcc = _build._nvcc()
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    csrc = os.path.join(_PKG, "csrc")
    defs = ["-DPMP_PLUGIN", f'-DPMP_GENERATED_HEADER="{hpath}"']

    def compile_one(src):
        obj = os.path.join(work, src.replace(".cu", ".o"))
        cmd = [nvcc, *_build.NVCC_FLAGS, *defs, "-ccbin", host, "-c", os.path.join(csrc, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"codegen: nvcc failed for {src} (generated header {hpath}):\n{r.stdout}\n{r.stderr}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(PLUGIN_SOURCES)) as ex:
        objs = list(ex.map(compile_one, PLUGIN_SOURCES))
    tmp = lib_path + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-ccbin", host, "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"codegen: link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib_path)
    return lib_path


_plugin_paths: dict = {}  # (f, l, nx, nu, U_interval, name) -> library path: repeated calls skip the sympy derivation


def plugin_problem(problem_params: dict) -> _capi.Problem:
    """pmp_problem_t for a generated system (system_id PMP_SYS_GENERATED) bound to its plug-in library.  Only the first call
    for a system pays the derivation (sympy trace, simplification, cse) and the compilation."""
    key = (problem_params.get("f"), problem_params.get("l"), int(problem_params["nx"]), int(problem_params["nu"]),
           repr(problem_params["U_interval"]), str(problem_params.get("system_name", "user")), problem_params.get("ustar_fct"))
    try:
        path = _plugin_paths.get(key)
    except TypeError:  # unhashable callables
        key, path = None, None
    if path is None or not os.path.exists(path):
        path = build_plugin(problem_params)
        if key is not None:
            _plugin_paths[key] = path
    nx, nu = int(problem_params["nx"]), int(problem_params["nu"])
    p = _capi.Problem(system_id=_capi.SYS_GENERATED, nx=nx, nu=nu)
    lo, hi = problem_params["U_interval"]
    lo = np.broadcast_to(np.asarray(lo, dtype=float), (nu,))
    hi = np.broadcast_to(np.asarray(hi, dtype=float), (nu,))
    for i in range(nu):
        p.u_lo[i], p.u_hi[i] = float(lo[i]), float(hi[i])
    p.lib_path = path
    _capi.load(path)  # fail now, not at the first solve
    return p
