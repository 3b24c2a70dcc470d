"""Device code for user-defined System subclasses.

The reference's contract for a new system is three methods, `_f_cont_fcn(x,u)`, `_l_fcn(x,u)`, `_l_f_fcn(x)`
(system_base.py:255-275); everything else -- the discrete step, f_x, f_u, l_x, l_u, l_xx, l_uu, l_ux, l_f_x,
l_f_xx -- JAX derives by tracing and autodiff (system_base.py:203-219).  Here the three methods are traced once
with sympy symbols (class_files/symbolic.py), differentiated analytically, passed through common-subexpression
elimination and printed as a CUDA header defining ilqr::UserSys<T> and ilqr::UserCost<T>.  The library's generic
kernel templates (csrc/*.cuh: integrators, chain rule through their stages, K1/K2/K3) are then compiled against that
header by NVRTC for sm_100a -- in this process, without nvcc or a host compiler, like the reference's jit at
construction -- and the cubin is handed to libilqr_b200.so (ilqr_module_load, ilqr_create_user; model ILQR_USER).
Cubins are cached under _user_cache/ by content hash.
"""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

from . import _cabi
from . import _nvrtc
from . import symbolic

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG)
CACHE = os.environ.get("ILQR_USER_CACHE") or os.path.join(_PKG, "_user_cache")
_FUNCS = {"sin": "sin_t", "cos": "cos_t", "tan": "tan_t", "exp": "exp_t", "log": "log_t", "sqrt": "sqrt_t",
          "tanh": "tanh_t", "Abs": "abs_t", "atan2": "atan2_t", "atan": "atan_t", "asin": "asin_t", "acos": "acos_t",
          "sinh": "sinh_t", "cosh": "cosh_t"}


class _Printer(C99CodePrinter):
    """C for `template <typename T>` device code: every literal is a T, every function its _t overload."""

    def __init__(self):
        super().__init__({"user_functions": dict(_FUNCS)})
        self.loop_local = {}                 # process-wide loop id (symbolic._ALL_LOOPS) -> number within this system

    def _loop(self, gid):
        return self.loop_local.setdefault(int(gid), len(self.loop_local))

    def _print_Symbol(self, e):
        m = re.fullmatch(r"wl(\d+)_([ctn])(\w+)", e.name)        # carry / tangent variables of a staged while_loop
        return f"wl{self._loop(m.group(1))}_{m.group(2)}{m.group(3)}" if m else super()._print_Symbol(e)

    def _print_LoopLeaf(self, e):
        return f"wl{self._loop(e.args[0])}_c{int(e.args[1])}"

    def _print_LoopTangent(self, e):
        return f"wl{self._loop(e.args[0])}_t{int(e.args[1])}_{int(e.args[2])}"

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


def _trace(system, loops=None):
    """the three user methods evaluated once on sympy symbols; `loops` (a dict) receives the staged lax.while_loops each
    method contains (symbolic._Loop), keyed "f" / "l" / "lf"""
    n, m = int(system.n_x), int(system.n_u)
    xs = [sp.Symbol(f"x[{i}]", real=True) for i in range(n)]
    us = [sp.Symbol(f"u[{j}]", real=True) for j in range(m)]
    x, u = np.array(xs, dtype=object), np.array(us, dtype=object)
    loops = {} if loops is None else loops
    try:
        symbolic.begin_trace(xs + us)
        f = np.asarray(system._f_cont_fcn(x, u), dtype=object).reshape(-1)
        loops["f"] = symbolic.end_trace()
        if f.shape[0] != n:
            raise ValueError(f"_f_cont_fcn must return {n} state derivatives, got {f.shape[0]}")
        f = [sp.sympify(v) for v in f]
        symbolic.begin_trace(xs + us)
        l = sp.sympify(np.asarray(system._l_fcn(x, u), dtype=object).reshape(-1)[0])
        loops["l"] = symbolic.end_trace()
        symbolic.begin_trace(xs)
        lf = sp.sympify(np.asarray(system._l_f_fcn(x), dtype=object).reshape(-1)[0])
        loops["lf"] = symbolic.end_trace()
    finally:
        symbolic.end_trace()
    return n, m, xs, us, f, l, lf


def _loop_code(pr, lp, tangents, indent):
    """One staged lax.while_loop as device code: the carry (and, for the Jacobian functions, its tangents with respect to
    the base symbols -- forward mode, as jax.jacfwd treats a while_loop) in function-scope variables, advanced by a real
    loop whose trip count is decided per trajectory by the traced condition."""
    g, base = lp.id, lp.base
    act = [i for i in range(len(lp.carry)) if lp.active[i]] if tangents else []
    tan = {(i, v): sp.Symbol(f"wl{g}_t{i}_{v}", real=True) for i in act for v in range(len(base)) if lp.nz[i][v]}
    nxt = [sp.Symbol(f"wl{g}_n{i}", real=True) for i in range(len(lp.carry))]
    ntan = {k: sp.Symbol(f"wl{g}_n{k[0]}_{k[1]}", real=True) for k in tan}
    names = [pr.doprint(c) for c in lp.carry] + [pr.doprint(t) for t in tan.values()]
    init = [(pr.doprint(c), e) for c, e in zip(lp.carry, lp.init)]
    init += [(pr.doprint(tan[i, v]), sp.diff(lp.init[i], base[v])) for (i, v) in tan]
    step = [(f"const T {pr.doprint(nxt[i])}", b) for i, b in enumerate(lp.body)]
    for (i, v) in tan:
        e = sp.diff(lp.body[i], base[v])
        for j in act:
            if (j, v) in tan:
                e += sp.diff(lp.body[i], lp.carry[j]) * tan[j, v]
        step.append((f"const T {pr.doprint(ntan[i, v])}", e))
    trip = f"wl{pr._loop(g)}_trip"
    inner = indent + "    "
    out = [f"{indent}T {', '.join(names)};",
           f"{indent}{{", _body(pr, init, inner), f"{indent}}}",
           f"{indent}int {trip} = 0;",
           f"{indent}for (; {trip} < ILQR_WHILE_MAX && ({pr.doprint(lp.cond)}); ++{trip}) {{",
           _body(pr, step, inner)]
    out += [f"{inner}{pr.doprint(c)} = {pr.doprint(nx)};" for c, nx in zip(lp.carry, nxt)]
    out += [f"{inner}{pr.doprint(tan[k])} = {pr.doprint(ntan[k])};" for k in tan]
    out += [f"{indent}}}",
            # a loop that never ends on the device would hang the GPU: past the cap the results are NaN, which no line
            # search accepts
            f"{indent}if ({trip} >= ILQR_WHILE_MAX) {{ " + " ".join(f"{nm} = T(__int_as_float(0x7fffffff));" for nm in names)
            + " }"]
    return "\n".join(out)


def _body(pr, assigns, indent="        ", loops=(), tangents=False):
    """C statements computing `assigns` = [(lhs, expr)] with shared subexpressions hoisted; `loops`: the staged
    while_loops the expressions read, emitted first"""
    exprs = [e for _, e in assigns]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("w_"), optimizations="basic") if exprs else ([], [])
    out = [_loop_code(pr, lp, tangents, indent) for lp in loops]
    out += [f"{indent}const T {pr.doprint(s)} = {pr.doprint(e)};" for s, e in repl]
    out += [f"{indent}{lhs} = {pr.doprint(e)};" for (lhs, _), e in zip(assigns, red)]
    return "\n".join(out)


def generate_header(system):
    """-> (header text, n, m).  See csrc/ilqr_systems.cuh for how UserSys/UserCost are used."""
    loops = {}
    n, m, xs, us, f, l, lf = _trace(system, loops)
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
    f_only = _body(pr, [(f"xd[{i}]", f[i]) for i in range(n)], loops=loops["f"])
    f_jac = _body(pr, [(f"xd[{i}]", f[i]) for i in range(n)] +
                  [(f"Ac[{i}][{j}]", Ac[i][j]) for i in range(n) for j in range(n)] +
                  [(f"Bc[{i}][{j}]", Bc[i][j]) for i in range(n) for j in range(m)], loops=loops["f"], tangents=True)
    # (a while_loop inside a cost would need second derivatives through the loop: symbolic.LoopTangent.fdiff raises)
    expand = _body(pr, [(f"lx[{i}]", lx[i]) for i in range(n)] + [(f"lu[{j}]", lu[j]) for j in range(m)] +
                   [(f"lxx[{i}][{j}]", lxx[i][j]) for i in range(n) for j in range(n)] +
                   [(f"luu[{i}][{j}]", luu[i][j]) for i in range(m) for j in range(m)] +
                   [(f"lux[{i}][{j}]", lux[i][j]) for i in range(m) for j in range(n)], loops=loops["l"], tangents=True)
    term = _body(pr, [(f"g[{i}]", lfx[i]) for i in range(n)] +
                 [(f"H[{i}][{j}]", lfxx[i][j]) for i in range(n) for j in range(n)], loops=loops["lf"], tangents=True)
    stage = _body(pr, [("const T value", l)], loops=loops["l"])
    terminal = _body(pr, [("const T value", lf)], loops=loops["lf"])
    name = type(system).__name__
    text = f"""// generated by class_files/codegen.py from {name}._f_cont_fcn/_l_fcn/_l_f_fcn -- do not edit
#pragma once
#ifndef ILQR_WHILE_MAX
#define ILQR_WHILE_MAX 65536      // trip cap of a generated lax.while_loop (results are NaN beyond it)
#endif
namespace ilqr {{
template <typename T>
struct UserSys {{
    static constexpr int NQ = 0, N = {n}, M = {m};
    static constexpr bool FIRST_ORDER = false, GENERIC = true, TRIG_TABLE = false;
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


def _sources():
    """{include name: text} of everything the kernel translation unit includes besides the generated model"""
    csrc = os.path.join(_PKG, "csrc")
    out = {name: open(os.path.join(csrc, name)).read() for name in sorted(os.listdir(csrc)) if name.endswith(".cuh")}
    out["ilqr_b200.h"] = open(os.path.join(_ROOT, "include", "ilqr_b200.h")).read()
    return out


# the translation unit NVRTC compiles: the generic kernel templates of the library against the generated model
_TU = """#include "ilqr_systems.cuh"
#include "ilqr_user.cuh"
#include "ilqr_kernels_common.cuh"
#include "ilqr_kernels_linearize.cuh"
#include "ilqr_kernels_backward.cuh"
#include "ilqr_kernels_rollout.cuh"
"""
_NVRTC_OPTIONS = ("--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo")


def kernel_expressions(n, m, integrator, dtype):
    """the template instantiations a module holds, in ILQR_UK_* order (include/ilqr_b200.h)"""
    T = "float" if dtype == "float32" else "double"
    S, Cst, I = f"ilqr::UserSys<{T}>", f"ilqr::UserCost<{T}>", _cabi.INTEGRATORS[integrator]
    small, large = ((2, 32), (2, 32)) if n > 4 else ((8, 32), (4, 64))
    return [f"ilqr::step_kernel<{S}, {I}, {T}>",
            f"ilqr::commit_linearize_kernel<{S}, {I}, {T}>",
            f"ilqr::cost_expansion_kernel<{Cst}, {T}, {n}, {m}>",
            f"ilqr::backward_kernel<{Cst}, {T}, {n}, {m}, {small[0]}, {small[1]}>",
            f"ilqr::backward_kernel<{Cst}, {T}, {n}, {m}, {large[0]}, {large[1]}>",
            f"ilqr::rollout_kernel<{S}, {Cst}, {I}, {T}>"]


def cubin_path(header_text, integrator, dtype):
    h = hashlib.sha1()
    for name, text in _sources().items():
        h.update(name.encode() + text.encode())
    h.update((header_text + integrator + dtype + _TU + " ".join(_NVRTC_OPTIONS) + "%d.%d" % _nvrtc.version()).encode())
    return os.path.join(CACHE, f"ilqr_user_{h.hexdigest()[:16]}.cubin")


def compile_module(system):
    """-> (cubin bytes, lowered kernel names in ILQR_UK_* order, n, m): generated model + kernel templates compiled by
    NVRTC in this process (no nvcc, no host compiler), or fetched from the on-disk cache keyed by the content of the
    model, the kernel sources and the NVRTC version."""
    text, n, m = generate_header(system)
    exprs = kernel_expressions(n, m, system.integrator, system.dtype)
    path = cubin_path(text, system.integrator, system.dtype)
    if os.path.exists(path) and os.path.exists(path + ".names"):
        names = open(path + ".names").read().split("\n")
        if len(names) == len(exprs):
            return open(path, "rb").read(), names, n, m
    headers = _sources()
    headers["ilqr_user.cuh"] = text
    try:
        cubin, lowered = _nvrtc.compile_cubin(_TU, headers, exprs, _NVRTC_OPTIONS)
    except RuntimeError as e:
        raise RuntimeError(f"NVRTC failed for the generated system {type(system).__name__}:\n{e}") from e
    names = [lowered[e] for e in exprs]
    try:                                             # the cache is an optimisation: a read-only tree still works
        os.makedirs(CACHE, exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as fh:
            fh.write(cubin)
        os.replace(tmp, path)
        with open(path + ".names", "w") as fh:
            fh.write("\n".join(names))
        with open(path[:-6] + ".cuh", "w") as fh:    # the generated model, for inspection
            fh.write(text)
    except OSError:
        pass
    return cubin, names, n, m


class UserLibrary:
    """libilqr_b200.so as seen by ONE user-defined system: every entry point of the C ABI, with ilqr_create bound to the
    module that holds the system's kernels (ilqr_module_load / ilqr_create_user)."""

    def __init__(self, system):
        self._lib = _cabi.load()
        cubin, names, n, m = compile_module(system)
        self._image = cubin                          # the runtime may reference the image while the module lives
        arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
        self._module = C.c_void_p()
        rc = self._lib.ilqr_module_load(cubin, len(cubin), arr, n, m, _cabi.INTEGRATORS[system.integrator],
                                        _cabi.DTYPES[system.dtype], C.byref(self._module))
        _cabi.check(rc)

    def ilqr_create(self, problem, out):
        return self._lib.ilqr_create_user(problem, self._module, out)

    def __getattr__(self, name):
        return getattr(self._lib, name)


_LIBRARIES = {}


def build_library(system):
    """the library object of a user-defined system (its kernels compiled by NVRTC on first use, then cached on disk;
    systems with the same generated code, integrator and element type share one loaded module per process)"""
    from . import _device
    _device.require_cuda()
    key = cubin_path(generate_header(system)[0], system.integrator, system.dtype)
    if key not in _LIBRARIES:
        _LIBRARIES[key] = UserLibrary(system)
    return _LIBRARIES[key]
