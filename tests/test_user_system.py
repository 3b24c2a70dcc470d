"""User-defined System subclasses: the reference's three-method contract
(`_f_cont_fcn`, `_l_fcn`, `_l_f_fcn`, system_base.py:255-275) through class_files.symbolic -> analytic
derivatives -> generated CUDA -> its own library with the same C ABI.

Golden vectors `tests/golden/user_cartpole_*.npz` come from the UNMODIFIED reference running the same class
body (tests/user_systems.py) with its own jit/autodiff factory (tests/golden/make_golden.py)."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_err
from helpers import UA_OL, cfg2_x0, golden_flow, ua_system
from user_systems import CARTPOLE, SPRING, make_cartpole_class, make_implicit_spring_class, make_user_ua_class

INTEGRATORS = ("euler", "midpoint", "rk4", "backward_euler")
TOL = 1e-9


def cartpole(integrator="rk4", dtype="float64"):
    from class_files import symbolic as jnp
    from class_files.systems.system_base import System
    p = CARTPOLE
    cls = make_cartpole_class(System, jnp)
    return cls(dt=p["dt"], x_target=np.array(p["x_target"]), Q=np.diag(p["Q"]), R=np.diag(p["R"]), Q_f=np.diag(p["Q_f"]),
               integrator=integrator, dtype=dtype, **{k: p[k] for k in ("mc", "mp", "l", "g", "b", "p_max", "w_bar")})


def user_ua(integrator="rk4"):
    from class_files import symbolic as jnp
    from class_files.systems.system_base import System
    p = UA_OL
    return make_user_ua_class(System, jnp)(dt=p["dt"], x_target=np.array(p["x_target"]), Q=np.diag(p["Q"]),
                                           R=np.diag(p["R"]), Q_f=np.diag(p["Q_f"]), integrator=integrator, **p["phys"])


# ------------------------------------------------------------------------------------------ CPU
def test_generated_header_derivatives_match_finite_differences():
    import sympy as sp
    from class_files import codegen
    s = cartpole()
    text, n, m = codegen.generate_header(s)
    assert (n, m) == (4, 1) and "struct UserSys" in text and "struct UserCost" in text
    assert "exp_t(" in text and "sin_t(" in text and "QUADRATIC = false" in text
    # the analytic continuous Jacobian that went into the header vs central differences of the traced f
    _, _, xs, us, f, l, lf = codegen._trace(s)
    fn = sp.lambdify([xs, us], f, "math")
    Ac = sp.lambdify([xs, us], [[sp.diff(fi, v) for v in xs] for fi in f], "math")
    lxx = sp.lambdify([xs, us], [[sp.diff(sp.diff(l, a), b) for b in xs] for a in xs], "math")
    lfn = sp.lambdify([xs, us], l, "math")
    rng = np.random.default_rng(0)
    for _ in range(5):
        x, u = rng.uniform(-1, 1, 4), rng.uniform(-2, 2, 1)
        A = np.array(Ac(x, u), dtype=float)
        H = np.array(lxx(x, u), dtype=float)
        for j in range(4):
            e = np.zeros(4); e[j] = 1e-6
            fd = (np.array(fn(x + e, u)) - np.array(fn(x - e, u))) / 2e-6
            assert np.allclose(A[:, j], fd, rtol=1e-6, atol=1e-7)
        e = np.zeros(4); e[0] = 1e-4
        fd2 = (lfn(x + e, u) - 2 * lfn(x, u) + lfn(x - e, u)) / 1e-8
        assert abs(H[0, 0] - fd2) < 1e-4 * max(1.0, abs(H[0, 0]))
    assert abs(H[0, 0] - CARTPOLE["Q"][0] * CARTPOLE["dt"]) > 1e-6        # the barrier makes l_xx state dependent


def test_user_kernels_compile_in_process_with_nvrtc(tmp_path, monkeypatch):
    """The kernels of a user-defined system are built by NVRTC inside this process (no nvcc, no host compiler, no GPU
    needed to compile): six instantiations of the library's generic templates, as a cubin for sm_100a, cached by
    content.  With PATH emptied, so that no toolchain binary could be found."""
    from class_files import codegen
    monkeypatch.setenv("PATH", "")
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    s = cartpole("midpoint")
    cubin, names, n, m = codegen.compile_module(s)
    assert (n, m) == (4, 1) and cubin[:4] == b"\x7fELF" and len(names) == 6
    for name, frag in zip(names, ("step_kernel", "commit_linearize_kernel", "cost_expansion_kernel", "backward_kernel",
                                  "backward_kernel", "rollout_kernel")):
        assert frag in name and "UserSys" in name or "UserCost" in name, name
    assert "Li1E" in names[0]                                     # midpoint = integrator 1 in the template arguments
    files = sorted(f.name for f in tmp_path.iterdir())
    assert len(files) == 3 and files[0].endswith(".cubin")       # cubin, lowered names, generated model
    again = codegen.compile_module(s)                             # second call: served from the cache
    assert again[0] == cubin and again[1] == names
    # another integrator or element type is another module
    assert codegen.cubin_path(codegen.generate_header(s)[0], "rk4", "float64") != codegen.cubin_path(
        codegen.generate_header(s)[0], "midpoint", "float64")


def _saturated(integ="rk4"):
    from class_files import symbolic
    from class_files.systems.system_base import System
    from user_systems import make_saturated_pendulum_class
    cls = make_saturated_pendulum_class(System, symbolic)
    return cls(dt=0.01, x_target=np.array([np.pi, 0.0]), Q=np.diag([1.0, 0.1]), R=np.diag([0.5]), Q_f=np.diag([50.0, 5.0]),
               integrator=integ)


def test_data_dependent_selects_are_traced_as_conditionals(tmp_path, monkeypatch):
    """jnp.where / clip in user methods: traced as Piecewise, differentiated branch by branch, generated as conditional
    expressions, compiled by NVRTC"""
    from class_files import codegen
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    s = _saturated()
    text, n, m = codegen.generate_header(s)
    assert (n, m) == (2, 1) and text.count("?") >= 4            # saturation in f and B_c, wall in l, l_x, l_xx
    cubin, names, _, _ = codegen.compile_module(s)
    assert cubin[:4] == b"\x7fELF"


def spring(integrator="rk4", dtype="float64"):
    from class_files import symbolic as jnp
    from class_files.systems.system_base import System
    p = SPRING
    cls = make_implicit_spring_class(System, jnp, jnp.lax)
    return cls(dt=p["dt"], x_target=np.array(p["x_target"]), Q=np.diag(p["Q"]), R=np.diag(p["R"]), Q_f=np.diag(p["Q_f"]),
               a=p["a"], ks=p["ks"], integrator=integrator, dtype=dtype)


def test_while_loop_is_staged_as_a_device_loop(tmp_path, monkeypatch):
    """lax.while_loop with a data-dependent trip count (the construct the reference's own integrator uses,
    system_base.py:139) inside a user's dynamics: one traced body, a real loop in the generated code that carries the
    forward-mode tangents beside the values, compiled by NVRTC; loop numbering is per system, so the text (and with it
    the cache key) is reproducible"""
    from class_files import codegen
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    s = spring()
    text, n, m = codegen.generate_header(s)
    assert (n, m) == (2, 1)
    f_only, f_jac = text.split("void f_jac")[0], text.split("void f_jac")[1].split("struct UserCost")[0]
    assert f_only.count("for (; wl0_trip < ILQR_WHILE_MAX") == 1 and "wl0_t0_0" not in f_only     # values only
    assert f_jac.count("for (; wl0_trip < ILQR_WHILE_MAX") == 1 and "wl0_t0_0 = wl0_n0_0;" in f_jac
    assert "wl0_t1_" not in text                                   # the trip counter does not depend on x, u: no tangents
    assert "wl0_t0_2" not in text and "wl0_t0_1" in text           # nor does the deflection depend on u (base symbol 2)
    assert "?" in f_jac                                            # the lax.cond became a select
    assert codegen.generate_header(spring())[0] == text            # a second trace prints the same text
    cubin, names, _, _ = codegen.compile_module(s)
    assert cubin[:4] == b"\x7fELF" and len(names) == 6


_HOST_SHIM = r"""
// host stand-ins for what the generated model uses from csrc/ilqr_systems.cuh (the device versions are PTX-backed)
#include <cmath>
#include <cstring>
#define ILQR_DEV inline
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
#define DEF1(name, fn) static inline double name(double x) { return fn(x); }
DEF1(sin_t, std::sin) DEF1(cos_t, std::cos) DEF1(tan_t, std::tan) DEF1(exp_t, std::exp) DEF1(log_t, std::log)
DEF1(sqrt_t, std::sqrt) DEF1(tanh_t, std::tanh) DEF1(abs_t, std::fabs) DEF1(atan_t, std::atan) DEF1(asin_t, std::asin)
DEF1(acos_t, std::acos) DEF1(sinh_t, std::sinh) DEF1(cosh_t, std::cosh)
static inline double pow_t(double x, double y) { return std::pow(x, y); }
static inline double atan2_t(double y, double x) { return std::atan2(y, x); }
#include "model.cuh"
extern "C" void f_jac(const double *x, const double *u, double *xd, double *Ac, double *Bc)
{
    constexpr int N = ilqr::UserSys<double>::N, M = ilqr::UserSys<double>::M;
    ilqr::UserSys<double> s;
    double A[N][N], B[N][M];
    s.f_jac(x, u, xd, A, B);
    std::memcpy(Ac, A, sizeof A);
    std::memcpy(Bc, B, sizeof B);
}
extern "C" void f_only(const double *x, const double *u, double *xd) { ilqr::UserSys<double> s; s.f(x, u, xd); }
extern "C" double cost_expand(double dt, const double *x, const double *u, double *lx, double *lu, double *lxx, double *luu,
                              double *lux, double *lf, double *lfx, double *lfxx)
{
    constexpr int N = ilqr::UserCost<double>::N, M = ilqr::UserCost<double>::M;
    ilqr::UserCost<double> c;
    c.dt = dt;
    double Lxx[N][N], Luu[M][M], Lux[M][N], H[N][N];
    c.expand(x, u, lx, lu, Lxx, Luu, Lux);
    c.terminal_expand(x, lfx, H);
    std::memcpy(lxx, Lxx, sizeof Lxx);
    std::memcpy(luu, Luu, sizeof Luu);
    std::memcpy(lux, Lux, sizeof Lux);
    std::memcpy(lfxx, H, sizeof H);
    *lf = c.terminal(x);
    return c.stage(x, u);
}
"""


def _host_model(system, tmp_path):
    """the generated model compiled for the HOST (g++), so that its code -- loops and tangents included -- runs on CPU"""
    import subprocess
    from class_files import codegen
    text, n, m = codegen.generate_header(system)
    (tmp_path / "model.cuh").write_text(text)
    (tmp_path / "host.cpp").write_text(_HOST_SHIM)
    so = tmp_path / "libmodel.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", str(tmp_path), "-o", str(so), str(tmp_path / "host.cpp")])
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)

    def f_jac(x, u):
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        xd, A, B = np.zeros(n), np.zeros((n, n)), np.zeros((n, m))
        lib.f_jac(x.ctypes.data_as(dp), u.ctypes.data_as(dp), xd.ctypes.data_as(dp), A.ctypes.data_as(dp), B.ctypes.data_as(dp))
        return xd, A, B

    def f(x, u):
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        xd = np.zeros(n)
        lib.f_only(x.ctypes.data_as(dp), u.ctypes.data_as(dp), xd.ctypes.data_as(dp))
        return xd
    def cost(x, u):
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        o = dict(lx=np.zeros(n), lu=np.zeros(m), lxx=np.zeros((n, n)), luu=np.zeros((m, m)), lux=np.zeros((m, n)),
                 lf=np.zeros(1), lfx=np.zeros(n), lfxx=np.zeros((n, n)))
        lib.cost_expand.restype = C.c_double
        o["l"] = lib.cost_expand(C.c_double(system.dt), x.ctypes.data_as(dp), u.ctypes.data_as(dp),
                                 *[o[k].ctypes.data_as(dp) for k in ("lx", "lu", "lxx", "luu", "lux", "lf", "lfx", "lfxx")])
        return o
    f.cost = cost
    return f, f_jac


def test_generated_cost_expansion_runs_on_the_host(tmp_path):
    """UserCost::stage / expand / terminal / terminal_expand of the cart-pole's barrier cost as GENERATED, compiled for the
    host: gradients and Hessians (state dependent: the barrier) against central differences of the generated values"""
    f, _ = _host_model(cartpole(), tmp_path)
    rng = np.random.default_rng(2)
    h = 1e-5
    for _ in range(10):
        x, u = rng.uniform(-1.2, 1.2, 4), rng.uniform(-2, 2, 1)
        o = f.cost(x, u)
        for j in range(4):
            e = np.zeros(4); e[j] = h
            op, om = f.cost(x + e, u), f.cost(x - e, u)
            assert abs(o["lx"][j] - (op["l"] - om["l"]) / (2 * h)) < 1e-7 * max(1.0, abs(o["lx"][j]))
            assert np.allclose(o["lxx"][:, j], (op["lx"] - om["lx"]) / (2 * h), rtol=1e-6, atol=1e-8)
            assert np.allclose(o["lux"][:, j], (op["lu"] - om["lu"]) / (2 * h), rtol=1e-6, atol=1e-8)
            assert abs(o["lfx"][j] - (op["lf"][0] - om["lf"][0]) / (2 * h)) < 1e-6 * max(1.0, abs(o["lfx"][j]))
            assert np.allclose(o["lfxx"][:, j], (op["lfx"] - om["lfx"]) / (2 * h), rtol=1e-6, atol=1e-7)
        up, um = f.cost(x, u + h), f.cost(x, u - h)
        assert abs(o["lu"][0] - (up["l"] - um["l"]) / (2 * h)) < 1e-8 and np.allclose(o["luu"][:, 0], (up["lu"] - um["lu"]) / (2 * h), atol=1e-8)
    xa, xb = np.array([0.2, 0.1, 0.0, 0.0]), np.array([1.4, 0.1, 0.0, 0.0])
    assert f.cost(xb, [0.0])["lxx"][0, 0] > 2 * f.cost(xa, [0.0])["lxx"][0, 0]          # the barrier's curvature near the wall


def test_generated_while_loop_code_runs_on_the_host(tmp_path):
    """The generated model of the implicit-spring pendulum compiled with g++ and run on the CPU: the staged loop stops where
    the data says, f and f_jac agree, the loop's forward-mode tangents equal the implicit-function-theorem derivative and
    central differences, and the lax.cond select switches with the sign of the velocity"""
    f, f_jac = _host_model(spring(), tmp_path)
    a, ks = SPRING["a"], SPRING["ks"]
    rng = np.random.default_rng(8)
    for _ in range(40):
        x, u = rng.uniform(-2.5, 2.5, 2), rng.uniform(-2, 2, 1)
        if abs(x[1]) < 1e-3:
            continue
        xd, A, B = f_jac(x, u)
        assert np.array_equal(xd, f(x, u))
        r = np.sin(x[0]) + 0.5 * x[1]
        y = r / (1 + a * r * r)
        for _ in range(60):
            y = y - (y + a * y ** 3 - r) / (1 + 3 * a * y * y)
        damp = 0.05 if x[1] > 0 else 0.15
        assert np.allclose(xd, [x[1], u[0] - damp * x[1] - 9.81 * np.sin(x[0]) - ks * y], rtol=1e-12, atol=1e-12)
        dy = 1.0 / (1 + 3 * a * y * y)                                       # implicit function theorem
        assert np.allclose(A, [[0.0, 1.0], [-9.81 * np.cos(x[0]) - ks * dy * np.cos(x[0]), -damp - ks * dy * 0.5]], rtol=1e-9, atol=1e-11)
        assert np.allclose(B[:, 0], [0.0, 1.0])
        for j in range(2):
            e = np.zeros(2); e[j] = 1e-6
            assert np.allclose(A[:, j], (f(x + e, u) - f(x - e, u)) / 2e-6, rtol=1e-6, atol=1e-7)


def test_two_sequential_while_loops_on_the_host(tmp_path):
    """a second loop that starts from, and keeps reading, the result of the first: the generated code numbers them 0 and 1,
    the second loop's tangents chain through the first one's; values against the closed forms, Jacobians against central
    differences"""
    from class_files import codegen, symbolic
    from class_files.systems.system_base import System
    from user_systems import make_two_loop_class
    s = make_two_loop_class(System, symbolic, symbolic.lax)()
    text = codegen.generate_header(s)[0]
    assert "wl0_trip" in text and "wl1_trip" in text and "wl2_trip" not in text and "wl0_t0_0" in text and "wl1_t0_0" in text
    assert codegen.compile_module(s)[0][:4] == b"\x7fELF"              # and the device build of the same text (NVRTC)
    f, f_jac = _host_model(s, tmp_path)
    rng = np.random.default_rng(15)
    for _ in range(30):
        x, u = rng.uniform(-2, 2, 2), rng.uniform(-1, 1, 1)
        xd, A, B = f_jac(x, u)
        sq = np.sqrt(1.5 + np.sin(x[0]) + 0.3 * u[0])
        z = 0.25 * sq
        for _ in range(400):
            z = 0.5 * np.cos(sq * z) + 0.1 * x[1]
        assert np.allclose(xd, [x[1], u[0] - sq * x[0] - (z * z + 2 * z + 3)], rtol=1e-11, atol=1e-12)
        for j in range(2):
            e = np.zeros(2); e[j] = 1e-6
            assert np.allclose(A[:, j], (f(x + e, u) - f(x - e, u)) / 2e-6, rtol=2e-6, atol=1e-7), (x, u, j)
        assert np.allclose(B[:, 0], (f(x, u + 1e-6) - f(x, u - 1e-6)) / 2e-6, rtol=2e-6, atol=1e-7)


def test_generated_code_of_the_wider_jnp_surface_runs_on_the_host(tmp_path):
    f, f_jac = _host_model(_rich(), tmp_path)
    rng = np.random.default_rng(4)
    xs, us = rng.uniform(-1.5, 1.5, (24, 3)), rng.uniform(-3, 3, (24, 1))
    xs = xs[np.abs(xs[:, 2]) > 1e-3]
    ref = _rich_numpy_f(xs, us[:len(xs)])
    for x, u, r in zip(xs, us, ref):
        xd, A, B = f_jac(x, u)
        assert np.allclose(xd, r, rtol=1e-12, atol=1e-13)
        for j in range(3):
            e = np.zeros(3); e[j] = 1e-6
            assert np.allclose(A[:, j], (f(x + e, u) - f(x - e, u)) / 2e-6, rtol=1e-6, atol=1e-7)
        assert np.allclose(B[:, 0], (f(x, u + 1e-6) - f(x, u - 1e-6)) / 2e-6, rtol=1e-6, atol=1e-7)


def test_lax_control_flow_is_traced():
    """cond / select / switch / fori_loop / scan of jax.lax on traced values; a while_loop inside a cost is refused
    with a message (its Hessian would need second derivatives through the loop)"""
    import sympy as sp
    from class_files import codegen, symbolic as jnp
    from class_files.systems.system_base import System
    lax = jnp.lax
    x = sp.Symbol("x", real=True)
    assert lax.fori_loop(0, 3, lambda i, v: v * x + i, 1.0) == ((1.0 * x + 0) * x + 1) * x + 2
    carry, ys = lax.scan(lambda c, e: (c + e * x, c), 0, jnp.array([1.0, 2.0, 3.0]))
    assert sp.simplify(carry - 6.0 * x) == 0 and ys.shape == (3,) and sp.simplify(ys[2] - 3.0 * x) == 0
    c2, (ya, yb) = lax.scan(lambda c, e: (c * e[0] + e[1], (c, 2 * c)), 1, (jnp.array([x, 2.0]), jnp.array([1.0, x])), reverse=True)
    assert sp.expand(c2 - ((1 * 2.0 + x) * x + 1.0)) == 0 and ya.shape == (2,) and ya[1] == 1 and sp.expand(yb[0] - 2 * (2.0 + x)) == 0
    both = lax.cond(x > 0, lambda v: (v, jnp.array([v, 2 * v])), lambda v: (-v, jnp.array([0.0, v])), x)
    assert both[0] == sp.Piecewise((x, x > 0), (-x, True)) and both[1].shape == (2,)
    assert lax.cond(True, lambda v: v + 1, lambda v: v - 1, x) == x + 1
    sw = lax.switch(sp.Symbol("k"), [lambda v: v, lambda v: 2 * v, lambda v: 3 * v], x)
    assert sw.subs(sp.Symbol("k"), 1) == 2 * x and sw.subs(sp.Symbol("k"), 0) == x and sw.subs(sp.Symbol("k"), 2) == 3 * x
    assert lax.select(x > 1, x, 1.0) == sp.Piecewise((x, x > 1), (1.0, True))

    class LoopInCost(System):
        def __init__(self):
            super().__init__(n_x=1, n_u=1, dt=0.01)

        def _f_cont_fcn(self, x, u):
            return jnp.array([u[0] - x[0]])

        def _l_fcn(self, x, u):
            y, _ = lax.while_loop(lambda c: (jnp.abs(c[0] * c[0] - x[0]) > 1e-12) & (c[1] < 9),
                                  lambda c: (0.5 * (c[0] + x[0] / c[0]), c[1] + 1), (x[0], 0))
            return y + u[0] * u[0]

        def _l_f_fcn(self, x):
            return x[0] * x[0]

    with pytest.raises(NotImplementedError, match="second derivatives through lax.while_loop"):
        codegen.generate_header(LoopInCost())


def _rich():
    from class_files import symbolic
    from class_files.systems.system_base import System
    from user_systems import make_rich_math_class
    return make_rich_math_class(System, symbolic)()


def _rich_numpy_f(xs, us):
    """the same method body evaluated with numpy as the array namespace"""
    from user_systems import make_rich_math_class
    cls = make_rich_math_class(object, np)
    obj = cls.__new__(cls)
    return np.stack([np.asarray(cls._f_cont_fcn(obj, x, u), dtype=float) for x, u in zip(xs, us)])


def test_wider_jnp_surface_is_generated(tmp_path, monkeypatch):
    """cross / linalg.det / trace / mean are traced away; arctan, arcsin, sinh, cosh reach the device as their _t overloads;
    sign is a select; the module compiles"""
    from class_files import codegen
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    s = _rich()
    text, n, m = codegen.generate_header(s)
    assert (n, m) == (3, 1)
    for frag in ("atan_t(", "asin_t(", "sinh_t(", "cosh_t(", "?"):
        assert frag in text, frag
    cubin, names, _, _ = codegen.compile_module(s)
    assert cubin[:4] == b"\x7fELF"


def test_shipped_systems_keep_their_device_models():
    s = ua_system()
    assert not s._is_user_defined() and s._device_model()[0] == "ua_double_pendulum"
    from class_files.systems.system_base import System

    class Incomplete(System):
        def _f_cont_fcn(self, x, u):
            return x
    with pytest.raises(NotImplementedError, match="no device model"):
        Incomplete(2, 1, 0.01)._device_model()


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("integ", INTEGRATORS)
def test_user_cartpole_point_functions_vs_reference(integ):
    g = load_golden(f"user_cartpole_derivs_{integ}")
    s = cartpole(integ)
    xs, us = g["xs"], g["us"]
    got = dict(f=s.f_fcn(xs, us), f_x=s.f_x_fcn(xs, us), f_u=s.f_u_fcn(xs, us), l=s.l_fcn(xs, us),
               l_x=s.l_x_fcn(xs, us), l_u=s.l_u_fcn(xs, us), l_xx=s.l_xx_fcn(xs, us), l_uu=s.l_uu_fcn(xs, us),
               l_ux=s.l_ux_fcn(xs, us), l_f=s.l_f_fcn(xs), l_f_x=s.l_f_x_fcn(xs), l_f_xx=s.l_f_xx_fcn(xs))
    for k, v in got.items():
        ref = g[k].reshape(np.shape(v))
        assert rel_err(v, ref, floor=1e-3) < 1e-11, (k, rel_err(v, ref, floor=1e-3))
    one = s.f_x_fcn(xs[0], us[0])                         # single point, reference shapes
    assert one.shape == (4, 4) and rel_err(one, g["f_x"][0]) < 1e-11


@pytest.mark.gpu
def test_user_cartpole_passes_vs_reference():
    from class_files.iLQR_class import iLQR
    g = load_golden("user_cartpole_passes_rk4")
    s = cartpole("rk4")
    N = int(g["N"])
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((1, N)), verbose=False)
    X_nom, U_nom, c0 = sol.forward_pass(g["x0"], 0.0, sol.X, g["U_nom"], sol.U_ff, sol.K)
    assert rel_err(X_nom, g["X_nom"]) < 1e-12 and rel_err(c0, g["cost0"]) < 1e-12
    U_ff, K = sol.backward_pass(g["X_nom"], g["U_nom"])
    assert rel_err(K, g["K"]) < TOL and rel_err(U_ff, g["U_ff"], floor=1e-6) < TOL
    for a in (1.0, 0.5, 0.125):
        tag = str(a).replace(".", "p")
        Xn, Un, c = sol.forward_pass(g["x0_b"], a, g["X_nom"], g["U_nom"], g["U_ff"], g["K"])
        assert rel_err(Xn, g[f"X_a{tag}"]) < TOL and rel_err(c, g[f"cost_a{tag}"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,integ", [("user_cartpole_solve_rk4_T1", "rk4"), ("user_cartpole_solve_be_T1", "backward_euler")])
def test_user_cartpole_solve_vs_reference(name, integ):
    """every iteration the reference executed, on identical inputs (1e-9), then the full solve's flow and result"""
    from class_files.iLQR_class import iLQR
    g = load_golden(name)
    s = cartpole(integ)
    N = int(g["N"])
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((1, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]), verbose=True)
    ref_idx, ref_costs = golden_flow(g)
    for i in range(len(ref_idx)):
        U_ff, K = sol.backward_pass(g["it_X"][i], g["it_U"][i])
        assert rel_err(K, g["it_K"][i]) < 1e-8, (i, rel_err(K, g["it_K"][i]))
        assert rel_err(U_ff, g["it_U_ff"][i], floor=1e-6) < 1e-8, i
        if ref_idx[i] < 0:
            continue
        Xn, Un, c = sol.forward_pass(g["x0"], 0.5 ** ref_idx[i], g["it_X"][i], g["it_U"][i], g["it_U_ff"][i], g["it_K"][i])
        X_ref = g["it_X"][i + 1] if i + 1 < len(ref_idx) else g["X"]
        assert rel_err(Xn, X_ref) < TOL and rel_err(c, ref_costs[i + 1]) < TOL, i
    X, U, cost = sol.optimize_trajectory()
    idx, alphas, costs = sol.trace(0)
    k = min(6, len(ref_idx), len(idx))
    assert np.array_equal(idx[:k], ref_idx[:k]), (idx, ref_idx)
    assert np.all(np.abs(costs[:k + 1] - ref_costs[:k + 1]) <= 1e-8 * np.abs(ref_costs[:k + 1]))
    if np.array_equal(idx, ref_idx):
        assert rel_err(cost, g["cost"]) < 1e-6 and rel_err(X, g["X"]) < 1e-4


@pytest.mark.gpu
def test_user_cartpole_mpc_vs_reference():
    """receding horizon with a user-defined optimizer model (rk4) and a user-defined plant (midpoint)"""
    from class_files.iLQR_class import iLQR
    from class_files.mpc import run_mpc
    g = load_golden("user_cartpole_mpc_T0p5")
    opt, plant = cartpole("rk4"), cartpole(str(g["p_integrator_plant"]))
    N, ticks = int(g["N"]), int(g["ticks"])
    sol = iLQR(opt, float(g["T"]), g["x0"], np.zeros((1, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]), verbose=False)
    r = run_mpc(sol, plant, g["x0"], ticks)
    assert list(r["iterations"]) == list(g["n_backward"])
    assert rel_err(r["X_sim"], g["X_sim"]) < 1e-7 and rel_err(r["costs"], g["costs"]) < 1e-7
    assert rel_err(r["U_sim"], g["U_sim"], floor=1e-3) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("integ", INTEGRATORS)
def test_user_ua_matches_reference_goldens_and_shipped_model(integ):
    """the UA double pendulum re-entered as a user system: generated code vs the reference's goldens for the
    shipped system and vs the hand-written device model"""
    g = load_golden(f"derivs_ua_{integ}")
    s, b = user_ua(integ), ua_system(integ)
    xs, us = g["xs"], g["us"]
    for k, fn in (("f", s.f_fcn), ("f_x", s.f_x_fcn), ("f_u", s.f_u_fcn), ("l_x", s.l_x_fcn), ("l_xx", s.l_xx_fcn),
                  ("l_uu", s.l_uu_fcn), ("l_ux", s.l_ux_fcn)):
        v = fn(xs, us)
        assert rel_err(v, g[k].reshape(np.shape(v)), floor=1e-3) < 1e-11, (k, integ)
    assert rel_err(s.f_x_fcn(xs, us), b.f_x_fcn(xs, us)) < 1e-12


@pytest.mark.gpu
def test_user_ua_batched_solve_matches_shipped_model(oracle):
    """the UA double pendulum re-entered as a USER-DEFINED system (generated device code): a batch solve, every member
    against the oracle of the shipped model (helpers.member_parity: X, U, K, U_ff, cost and control flow)"""
    from class_files.iLQR_class import iLQR
    from helpers import member_parity, gpu_result, ua_oracle_problem, write_report
    B, N = 256, 100
    x0 = cfg2_x0(B, seed=4)
    sol = iLQR(user_ua(), 1.0, x0, np.zeros((1, N)), maxiter=3, verbose=False)
    sol.enable_trace()
    X, U, cost = sol.optimize_trajectory()
    report, failures = member_parity(oracle, ua_oracle_problem(oracle, N, maxiter=3), x0, np.zeros((B, 1, N)),
                                     gpu_result(sol, X, U, cost))
    write_report("user_defined_ua_B256_N100_it3", report)
    assert not failures, (failures[:5], report)


@pytest.mark.gpu
def test_never_seen_user_system_builds_on_this_machine_without_nvcc(oracle, tmp_path, monkeypatch):
    """A system nobody has compiled before -- a random constant is baked into its dynamics, so its content hash is new
    and the cache (redirected to an empty directory) cannot serve it -- is traced, differentiated, compiled by NVRTC and
    loaded on THIS machine with PATH emptied (no nvcc reachable), then checked against central finite differences of
    its own step and, its dynamics being the UA double pendulum's up to the random damping, against the oracle."""
    import secrets
    from class_files import codegen, symbolic
    from class_files.iLQR_class import iLQR
    from class_files.systems.system_base import System
    from helpers import UA_OL, ua_oracle_problem
    from user_systems import make_user_ua_class
    monkeypatch.setenv("PATH", "")
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    d1 = 0.05 + (secrets.randbits(40) / 2.0**40) * 0.1            # never the same twice
    phys = dict(UA_OL["phys"], d1=d1)
    cls = make_user_ua_class(System, symbolic)
    s = cls(dt=0.01, x_target=np.array(UA_OL["x_target"]), Q=np.diag(UA_OL["Q"]), R=np.diag(UA_OL["R"]),
            Q_f=np.diag(UA_OL["Q_f"]), integrator="rk4", **phys)
    assert not any(f.name.endswith(".cubin") for f in tmp_path.iterdir())
    rng = np.random.default_rng(5)
    xs, us = rng.uniform(-1, 1, (16, 4)), rng.uniform(-1, 1, (16, 1))
    A, Bm = s.f_x_fcn(xs, us), s.f_u_fcn(xs, us)                  # first use: trace -> NVRTC -> module load
    assert sum(f.name.endswith(".cubin") for f in tmp_path.iterdir()) == 1
    for j in range(4):
        e = np.zeros(4); e[j] = 1e-6
        fd = (s.f_fcn(xs + e, us) - s.f_fcn(xs - e, us)) / 2e-6
        assert np.allclose(A[:, :, j], fd, rtol=1e-6, atol=1e-8)
    fd = (s.f_fcn(xs, us + 1e-6) - s.f_fcn(xs, us - 1e-6)) / 2e-6
    assert np.allclose(Bm[:, :, 0], fd, rtol=1e-6, atol=1e-8)
    # a batched solve against the oracle of the same physical parameters
    B, N = 64, 60
    x0 = cfg2_x0(B, seed=6)
    sol = iLQR(s, N * 0.01, x0, np.zeros((1, N)), maxiter=3, verbose=False)
    sol.enable_trace()
    X, U, cost = sol.optimize_trajectory()
    from helpers import member_parity, gpu_result
    p = oracle.make_problem("ua", "rk4", N, 0.01, UA_OL["Q"], UA_OL["R"], UA_OL["Q_f"], UA_OL["x_target"], phys, maxiter=3)
    report, failures = member_parity(oracle, p, x0, np.zeros((B, 1, N)), gpu_result(sol, X, U, cost))
    assert not failures, (failures[:5], report)


@pytest.mark.gpu
def test_user_system_with_data_dependent_selects_vs_finite_differences():
    """saturated torque (clip) and a one-sided cost wall (where): every branch of the generated derivatives against
    central finite differences of the generated step / cost, and a batched solve that improves every member"""
    from class_files.iLQR_class import iLQR
    s = _saturated()
    rng = np.random.default_rng(9)
    xs = rng.uniform(-2.5, 2.5, (64, 2))
    us = rng.uniform(-4.0, 4.0, (64, 1))                              # half of them beyond the saturation
    keep = (np.abs(np.abs(us[:, 0]) - 2.0) > 1e-3) & (np.abs(xs[:, 1] - 1.0) > 1e-3)   # away from the kinks
    xs, us = xs[keep], us[keep]
    assert (np.abs(us[:, 0]) > 2).any() and (np.abs(us[:, 0]) < 2).any() and (xs[:, 1] > 1).any() and (xs[:, 1] < 1).any()
    A, Bm, lx, lxx = s.f_x_fcn(xs, us), s.f_u_fcn(xs, us), s.l_x_fcn(xs, us), s.l_xx_fcn(xs, us)
    h = 1e-6
    for j in range(2):
        e = np.zeros(2); e[j] = h
        assert np.allclose(A[:, :, j], (s.f_fcn(xs + e, us) - s.f_fcn(xs - e, us)) / (2 * h), rtol=1e-6, atol=1e-8)
        assert np.allclose(lx[:, j], (s.l_fcn(xs + e, us) - s.l_fcn(xs - e, us)) / (2 * h), rtol=1e-5, atol=1e-9)
        assert np.allclose(lxx[:, :, j], (s.l_x_fcn(xs + e, us) - s.l_x_fcn(xs - e, us)) / (2 * h), rtol=1e-5, atol=1e-9)
    fd = (s.f_fcn(xs, us + h) - s.f_fcn(xs, us - h)) / (2 * h)
    assert np.allclose(Bm[:, :, 0], fd, rtol=1e-6, atol=1e-8)
    assert np.all(Bm[np.abs(us[:, 0]) > 2, :, 0] == 0.0)                # saturated: the control has no effect
    B, N = 128, 80
    x0 = rng.uniform(-1.0, 1.0, (B, 2))
    sol = iLQR(s, N * 0.01, x0, np.zeros((1, N)), maxiter=10, verbose=False)
    c0 = sol.forward_pass(x0, 0.0, np.zeros((B, 2, N + 1)), np.zeros((B, 1, N)), np.zeros((B, 1, N)), np.zeros((B, N, 1, 2)))[2]
    X, U, cost = sol.optimize_trajectory()
    assert np.all(cost <= c0) and np.all(np.isfinite(cost))


@pytest.mark.gpu
def test_user_cartpole_batch_properties_and_fp32():
    import torch
    from class_files.iLQR_class import iLQR
    B, N = 2048, 150
    rng = np.random.default_rng(7)
    x0 = rng.uniform(-0.5, 0.5, (B, 4))
    sol = iLQR(cartpole(), 1.5, torch.as_tensor(x0).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"),
               maxiter=15, verbose=False)
    X, U, cost = sol.optimize_trajectory()
    zX, zU = torch.zeros_like(X), torch.zeros_like(U)
    c0 = sol.forward_pass(sol.x_0, 0.0, zX, zU, torch.zeros_like(sol.U_ff), torch.zeros_like(sol.K))[2]
    assert bool((cost <= c0).all()) and float((cost < 0.9 * c0).double().mean()) > 0.9
    Xr, Ur, cr = sol.forward_pass(sol.x_0, 0.0, X, U, torch.zeros_like(sol.U_ff), torch.zeros_like(sol.K))
    assert torch.equal(Xr, X) and torch.equal(cr, cost)
    # FP32 build of the same user system (1e-4 mode of BASELINE.json)
    s32 = cartpole(dtype="float32")
    g = load_golden("user_cartpole_derivs_rk4")
    assert rel_err(s32.f_x_fcn(g["xs"], g["us"]), g["f_x"], floor=1e-2) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("integ", INTEGRATORS)
def test_user_while_loop_point_functions_vs_reference(integ):
    """the implicit-spring pendulum (Newton iteration in a lax.while_loop + a lax.cond in _f_cont_fcn) against the
    UNMODIFIED reference running the same class body: its jacfwd differentiates through the loop (system_base.py:204-205),
    the generated device loop carries the same tangents"""
    g = load_golden(f"user_spring_derivs_{integ}")
    s = spring(integ)
    xs, us = g["xs"], g["us"]
    got = dict(f=s.f_fcn(xs, us), f_x=s.f_x_fcn(xs, us), f_u=s.f_u_fcn(xs, us), l=s.l_fcn(xs, us),
               l_x=s.l_x_fcn(xs, us), l_xx=s.l_xx_fcn(xs, us), l_f_x=s.l_f_x_fcn(xs))
    for k, v in got.items():
        ref = g[k].reshape(np.shape(v))
        assert rel_err(v, ref, floor=1e-3) < 1e-11, (k, rel_err(v, ref, floor=1e-3))
    # and the implicit function theorem: d y / d r = 1 / (1 + 3 a y^2) at the converged deflection
    if integ == "euler":
        a, ks, dt = SPRING["a"], SPRING["ks"], SPRING["dt"]
        r = np.sin(xs[:, 0]) + 0.5 * xs[:, 1]
        y = r / (1 + a * r * r)
        for _ in range(60):
            y = y - (y + a * y ** 3 - r) / (1 + 3 * a * y * y)
        damp = np.where(xs[:, 1] > 0, 0.05, 0.15)
        A10 = dt * (-9.81 * np.cos(xs[:, 0]) - ks * np.cos(xs[:, 0]) / (1 + 3 * a * y * y))
        A11 = 1 + dt * (-damp - ks * 0.5 / (1 + 3 * a * y * y))
        assert np.allclose(got["f_x"][:, 1, 0], A10, rtol=1e-10, atol=1e-12)
        assert np.allclose(got["f_x"][:, 1, 1], A11, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_user_while_loop_passes_vs_reference():
    from class_files.iLQR_class import iLQR
    g = load_golden("user_spring_passes_rk4")
    s = spring("rk4")
    N = int(g["N"])
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((1, N)), verbose=False)
    X_nom, U_nom, c0 = sol.forward_pass(g["x0"], 0.0, sol.X, g["U_nom"], sol.U_ff, sol.K)
    assert rel_err(X_nom, g["X_nom"]) < 1e-12 and rel_err(c0, g["cost0"]) < 1e-12
    U_ff, K = sol.backward_pass(g["X_nom"], g["U_nom"])
    assert rel_err(K, g["K"]) < TOL and rel_err(U_ff, g["U_ff"], floor=1e-6) < TOL
    for a in (1.0, 0.5, 0.125):
        tag = str(a).replace(".", "p")
        Xn, Un, c = sol.forward_pass(g["x0_b"], a, g["X_nom"], g["U_nom"], g["U_ff"], g["K"])
        assert rel_err(Xn, g[f"X_a{tag}"]) < TOL and rel_err(c, g[f"cost_a{tag}"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,integ", [("user_spring_solve_rk4_T1", "rk4"), ("user_spring_solve_midpoint_T1", "midpoint")])
def test_user_while_loop_solve_vs_reference(name, integ):
    """every iteration the reference executed, on identical inputs (1e-9), then the full solve's flow and result; and a
    batch whose members need different trip counts per step"""
    from class_files.iLQR_class import iLQR
    g = load_golden(name)
    s = spring(integ)
    N = int(g["N"])
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((1, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]), verbose=False)
    sol.enable_trace()
    ref_idx, ref_costs = golden_flow(g)
    for i in range(len(ref_idx)):
        U_ff, K = sol.backward_pass(g["it_X"][i], g["it_U"][i])
        assert rel_err(K, g["it_K"][i]) < 1e-8, (i, rel_err(K, g["it_K"][i]))
        assert rel_err(U_ff, g["it_U_ff"][i], floor=1e-6) < 1e-8, i
        if ref_idx[i] < 0:
            continue
        Xn, Un, c = sol.forward_pass(g["x0"], 0.5 ** ref_idx[i], g["it_X"][i], g["it_U"][i], g["it_U_ff"][i], g["it_K"][i])
        X_ref = g["it_X"][i + 1] if i + 1 < len(ref_idx) else g["X"]
        assert rel_err(Xn, X_ref) < TOL and rel_err(c, ref_costs[i + 1]) < TOL, i
    X, U, cost = sol.optimize_trajectory()
    idx, alphas, costs = sol.trace(0)
    k = min(6, len(ref_idx), len(idx))
    assert np.array_equal(idx[:k], ref_idx[:k]), (idx, ref_idx)
    assert np.all(np.abs(costs[:k + 1] - ref_costs[:k + 1]) <= 1e-8 * np.abs(ref_costs[:k + 1]))
    if np.array_equal(idx, ref_idx):
        assert rel_err(cost, g["cost"]) < 1e-6 and rel_err(X, g["X"]) < 1e-4
    # batched: member 0 is the golden's initial state, the others start elsewhere (other trip counts in the same warp)
    B = 96
    rng = np.random.default_rng(12)
    x0 = np.concatenate([np.asarray(g["x0"]).reshape(1, 2), rng.uniform(-2.0, 2.0, (B - 1, 2))])
    solb = iLQR(s, float(g["T"]), x0, np.zeros((1, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]), verbose=False)
    Xb, Ub, cb = solb.optimize_trajectory()
    assert np.all(np.isfinite(cb))
    assert rel_err(cb[0], cost) < 1e-12 and rel_err(Xb[0], X) < 1e-12       # a member does not depend on its neighbours


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["cartpole", "spring"])
def test_user_system_backward_bulk_copy_ring_is_exact(monkeypatch, which):
    """the NVRTC-compiled generic Riccati kernel of a user system (state-dependent cost Hessians): its bulk-copy ring
    (cp.async.bulk + mbarrier, the default from 16384 trajectories up, forced here) against the per-thread cp.async ring,
    bit for bit, over a solve with trajectories finishing at different iterations"""
    from class_files.iLQR_class import iLQR
    s = cartpole() if which == "cartpole" else spring()
    B, N = 512, 100
    rng = np.random.default_rng(21)
    x0 = rng.uniform(-0.5, 0.5, (B, s.n_x))
    out = {}
    for bulk in ("0", "1"):
        monkeypatch.setenv("ILQR_BACKWARD_BULK", bulk)
        sol = iLQR(s, N * 0.01, x0, np.zeros((1, N)), tol=1e-3, maxiter=25, verbose=False)
        X, U, cost = sol.optimize_trajectory()
        U_ff, K = sol.backward_pass(X, U)
        out[bulk] = [np.array(a) for a in (X, U, cost, sol.K, sol.U_ff, sol.iterations, sol.status, U_ff, K)]
    assert len(np.unique(out["0"][5])) > 1
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.gpu
def test_wider_jnp_surface_on_device():
    """the generated continuous dynamics (euler: f = x + dt f_c) against the SAME method body run with numpy, and the
    Jacobians against central finite differences"""
    from class_files import symbolic
    from class_files.systems.system_base import System
    from user_systems import make_rich_math_class
    s = make_rich_math_class(System, symbolic)(integrator="euler")
    rng = np.random.default_rng(4)
    xs, us = rng.uniform(-1.5, 1.5, (48, 3)), rng.uniform(-3, 3, (48, 1))
    xs = xs[np.abs(xs[:, 2]) > 1e-3]                               # away from the kink of sign()
    us = us[:len(xs)]
    fc = (s.f_fcn(xs, us) - xs) / 0.01
    assert np.allclose(fc, _rich_numpy_f(xs, us), rtol=1e-9, atol=1e-11)
    A, Bm = s.f_x_fcn(xs, us), s.f_u_fcn(xs, us)
    h = 1e-6
    for j in range(3):
        e = np.zeros(3); e[j] = h
        assert np.allclose(A[:, :, j], (s.f_fcn(xs + e, us) - s.f_fcn(xs - e, us)) / (2 * h), rtol=1e-6, atol=1e-8)
    assert np.allclose(Bm[:, :, 0], (s.f_fcn(xs, us + h) - s.f_fcn(xs, us - h)) / (2 * h), rtol=1e-6, atol=1e-8)


@pytest.mark.gpu
def test_two_sequential_while_loops_on_device(tmp_path):
    """the same model on the GPU (euler step: f = x + dt f_c) against its host build, point by point, and a batched solve"""
    from class_files import symbolic
    from class_files.iLQR_class import iLQR
    from class_files.systems.system_base import System
    from user_systems import make_two_loop_class
    cls = make_two_loop_class(System, symbolic, symbolic.lax)
    s = cls(integrator="euler")
    f, f_jac = _host_model(s, tmp_path)
    rng = np.random.default_rng(16)
    xs, us = rng.uniform(-2, 2, (40, 2)), rng.uniform(-1, 1, (40, 1))
    fx, A, Bm = s.f_fcn(xs, us), s.f_x_fcn(xs, us), s.f_u_fcn(xs, us)
    for i in range(len(xs)):
        xd, Ac, Bc = f_jac(xs[i], us[i])
        assert np.allclose(fx[i], xs[i] + 0.01 * xd, rtol=1e-12, atol=1e-13)
        assert np.allclose(A[i], np.eye(2) + 0.01 * Ac, rtol=1e-10, atol=1e-12)
        assert np.allclose(Bm[i], 0.01 * Bc, rtol=1e-10, atol=1e-13)
    B, N = 64, 60
    x0 = rng.uniform(-1, 1, (B, 2))
    sol = iLQR(cls(integrator="rk4"), N * 0.01, x0, np.zeros((1, N)), maxiter=8, verbose=False)
    c0 = sol.forward_pass(x0, 0.0, np.zeros((B, 2, N + 1)), np.zeros((B, 1, N)), np.zeros((B, 1, N)), np.zeros((B, N, 1, 2)))[2]
    X, U, cost = sol.optimize_trajectory()
    assert np.all(np.isfinite(cost)) and np.all(cost <= c0) and np.mean(cost < c0) > 0.9
