"""Device code for user-defined System subclasses.

The reference's contract for a new system is three methods, `_f_cont_fcn(x,u)`, `_l_fcn(x,u)`, `_l_f_fcn(x)`
(system_base.py:255-275); everything else -- the discrete step, f_x, f_u, l_x, l_u, l_xx, l_uu, l_ux, l_f_x,
l_f_xx -- JAX derives by tracing and autodiff (system_base.py:203-219).  Here the three methods are traced once
with sympy symbols (class_files/symbolic.py), differentiated analytically, passed through common-subexpression
elimination and printed as a CUDA header defining ilqr::UserSys<T> and ilqr::UserCost<T>.  csrc/ilqr_b200.cu is
then compiled against that header with nvcc for sm_100a into its own shared library with the SAME C ABI
(include/ilqr_b200.h, model ILQR_USER), cached in-tree under _user_cache/ by content hash.  The integrators and
the chain rule through their stages are the generic device templates of csrc/ilqr_systems.cuh.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

from . import _cabi

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG)
CACHE = os.environ.get("ILQR_USER_CACHE") or os.path.join(_PKG, "_user_cache")
_FUNCS = {"sin": "sin_t", "cos": "cos_t", "tan": "tan_t", "exp": "exp_t", "log": "log_t", "sqrt": "sqrt_t",
          "tanh": "tanh_t", "Abs": "abs_t", "atan2": "atan2_t"}


class _Printer(C99CodePrinter):
    """C for `template <typename T>` device code: every literal is a T, every function its _t overload."""

    def __init__(self):
        super().__init__({"user_functions": dict(_FUNCS)})

    def _print_Float(self, e):
        return f"T({float(e)!r})"

    def _print_Integer(self, e):
        return f"T({int(e)})"

    def _print_Rational(self, e):
        return f"(T({int(e.p)}) / T({int(e.q)}))"

    def _print_Pow(self, e):
        b, p = e.base, e.exp
        pb = self.parenthesize(b, 1000)
        if p.is_Integer and 2 <= int(p) <= 4:
            return "(" + " * ".join([pb] * int(p)) + ")"
        if p.is_Integer and -4 <= int(p) <= -1:
            return "(T(1) / (" + " * ".join([pb] * -int(p)) + "))"
        if p == sp.Rational(1, 2):
            return f"sqrt_t({self._print(b)})"
        if p == sp.Rational(-1, 2):
            return f"(T(1) / sqrt_t({self._print(b)}))"
        return f"pow_t({self._print(b)}, {self._print(p)})"


def _trace(system):
    n, m = int(system.n_x), int(system.n_u)
    xs = [sp.Symbol(f"x[{i}]", real=True) for i in range(n)]
    us = [sp.Symbol(f"u[{j}]", real=True) for j in range(m)]
    x, u = np.array(xs, dtype=object), np.array(us, dtype=object)
    f = np.asarray(system._f_cont_fcn(x, u), dtype=object).reshape(-1)
    if f.shape[0] != n:
        raise ValueError(f"_f_cont_fcn must return {n} state derivatives, got {f.shape[0]}")
    f = [sp.sympify(v) for v in f]
    l = sp.sympify(np.asarray(system._l_fcn(x, u), dtype=object).reshape(-1)[0])
    lf = sp.sympify(np.asarray(system._l_f_fcn(x), dtype=object).reshape(-1)[0])
    return n, m, xs, us, f, l, lf


def _body(pr, assigns, indent="        "):
    """C statements computing `assigns` = [(lhs, expr)] with shared subexpressions hoisted"""
    exprs = [e for _, e in assigns]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("w_"), optimizations="basic") if exprs else ([], [])
    out = [f"{indent}const T {pr.doprint(s)} = {pr.doprint(e)};" for s, e in repl]
    out += [f"{indent}{lhs} = {pr.doprint(e)};" for (lhs, _), e in zip(assigns, red)]
    return "\n".join(out)


def generate_header(system):
    """-> (header text, n, m).  See csrc/ilqr_systems.cuh for how UserSys/UserCost are used."""
    n, m, xs, us, f, l, lf = _trace(system)
    pr = _Printer()
    Ac = [[sp.diff(f[i], xs[j]) for j in range(n)] for i in range(n)]
    Bc = [[sp.diff(f[i], us[j]) for j in range(m)] for i in range(n)]
    lx = [sp.diff(l, v) for v in xs]
    lu = [sp.diff(l, v) for v in us]
    lxx = [[sp.diff(lx[i], xs[j]) for j in range(n)] for i in range(n)]
    luu = [[sp.diff(lu[i], us[j]) for j in range(m)] for i in range(m)]
    lux = [[sp.diff(lu[i], xs[j]) for j in range(n)] for i in range(m)]          # jacfwd(grad_u l)_x, system_base.py:216
    lfx = [sp.diff(lf, v) for v in xs]
    lfxx = [[sp.diff(lfx[i], xs[j]) for j in range(n)] for i in range(n)]
    f_only = _body(pr, [(f"xd[{i}]", f[i]) for i in range(n)])
    f_jac = _body(pr, [(f"xd[{i}]", f[i]) for i in range(n)] +
                  [(f"Ac[{i}][{j}]", Ac[i][j]) for i in range(n) for j in range(n)] +
                  [(f"Bc[{i}][{j}]", Bc[i][j]) for i in range(n) for j in range(m)])
    expand = _body(pr, [(f"lx[{i}]", lx[i]) for i in range(n)] + [(f"lu[{j}]", lu[j]) for j in range(m)] +
                   [(f"lxx[{i}][{j}]", lxx[i][j]) for i in range(n) for j in range(n)] +
                   [(f"luu[{i}][{j}]", luu[i][j]) for i in range(m) for j in range(m)] +
                   [(f"lux[{i}][{j}]", lux[i][j]) for i in range(m) for j in range(n)])
    term = _body(pr, [(f"g[{i}]", lfx[i]) for i in range(n)] +
                 [(f"H[{i}][{j}]", lfxx[i][j]) for i in range(n) for j in range(n)])
    stage = _body(pr, [("const T value", l)])
    terminal = _body(pr, [("const T value", lf)])
    name = type(system).__name__
    text = f"""// generated by class_files/codegen.py from {name}._f_cont_fcn/_l_fcn/_l_f_fcn -- do not edit
#pragma once
namespace ilqr {{
template <typename T>
struct UserSys {{
    static constexpr int NQ = 0, N = {n}, M = {m};
    static constexpr bool FIRST_ORDER = false, GENERIC = true;
    ILQR_DEV T time_scalar(int, T) const {{ return T(0); }}
    ILQR_DEV void f(const T *x, const T *u, T *xd) const
    {{
{f_only}
    }}
    ILQR_DEV void f_jac(const T *x, const T *u, T *xd, T (*Ac)[N], T (*Bc)[M]) const
    {{
{f_jac}
    }}
}};
template <typename T>
struct UserCost {{
    static constexpr bool QUADRATIC = false;
    static constexpr int N = {n}, M = {m};
    T dt;
    int diag = 0, monotone = 0;
    ILQR_DEV T stage(const T *x, const T *u) const
    {{
{stage}
        return value;
    }}
    ILQR_DEV T terminal(const T *x) const
    {{
{terminal}
        return value;
    }}
    ILQR_DEV void expand(const T *x, const T *u, T *lx, T *lu, T (*lxx)[N], T (*luu)[M], T (*lux)[N]) const
    {{
{expand}
    }}
    ILQR_DEV void terminal_expand(const T *x, T *g, T (*H)[N]) const
    {{
{term}
    }}
}};
}}  // namespace ilqr
"""
    return text, n, m


def _sources_digest():
    h = hashlib.sha1()
    csrc = os.path.join(_PKG, "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(csrc, name), "rb").read())
    h.update(open(os.path.join(_ROOT, "include", "ilqr_b200.h"), "rb").read())
    return h.hexdigest()


def library_path(header_text, integrator, dtype):
    key = hashlib.sha1((header_text + integrator + dtype + _sources_digest()).encode()).hexdigest()[:16]
    return os.path.join(CACHE, f"libilqr_user_{key}.so"), os.path.join(CACHE, f"ilqr_user_{key}.cuh")


def build_library(system):
    """Generate + compile (or fetch from the in-tree cache) the library of `system`; returns the loaded CDLL."""
    text, n, m = generate_header(system)
    so, hdr = library_path(text, system.integrator, system.dtype)
    if not os.path.exists(so):
        os.makedirs(CACHE, exist_ok=True)
        with open(hdr, "w") as fh:
            fh.write(text)
        from importlib import util
        spec = util.spec_from_file_location("ilqr_b200_build", os.path.join(_PKG, "build.py"))
        bld = util.module_from_spec(spec)
        spec.loader.exec_module(bld)
        tmp = so + f".tmp{os.getpid()}"
        cmd = [bld.nvcc_path()] + bld.NVCC_FLAGS + [
            "-DILQR_USER_SYS", f'-DILQR_USER_HEADER="{hdr}"', f"-DILQR_USER_INTEG={_cabi.INTEGRATORS[system.integrator]}",
            f"-DILQR_USER_F32={1 if system.dtype == 'float32' else 0}", "-o", tmp] + bld.SRC
        try:
            subprocess.run(cmd, check=True, capture_output=True, text=True)
        except subprocess.CalledProcessError as e:
            raise RuntimeError(f"nvcc failed for the generated system {type(system).__name__}:\n{e.stderr[-4000:]}") from e
        os.replace(tmp, so)
    lib = C.CDLL(so)
    for fname, (res, args) in _cabi.SIGNATURES.items():
        fn = getattr(lib, fname)
        fn.restype, fn.argtypes = res, args
    return lib
