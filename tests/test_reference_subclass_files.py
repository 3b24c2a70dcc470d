"""Drop-in proof for user-defined systems: the reference's OWN System subclass files -- written with `import jax.numpy as
jnp`, never edited, read from /root/reference where it is mounted -- load against this package's System base class once
the opt-in alias directory (iterative-linear-quadratic-regulator_b200/jax_alias) is on sys.path, are traced, differentiated,
turned into device code and compiled by NVRTC in process (no GPU needed to compile).  The traced dynamics and their
analytic Jacobians are evaluated numerically and compared with the oracle's restatement of the same files.
Skipped where the reference is not mounted (the GPU box)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

REF = "/root/reference/python/class_files/systems"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources are not mounted here")

PHYS = dict(g=9.81, m1=1.0, m2=1.2, l1=1.0, l2=0.8, d1=0.1, d2=0.05, theta1=1.0 / 12, theta2=0.07)
CASES = {
    "pendulum": ("pendulum_sys.py", "MyPendulum", 2, 1, dict(g=9.81, l=0.9, d=0.05)),
    "double": ("double_pendulum_sys.py", "MyDoublePendulum", 4, 2, PHYS),
    "ua": ("UA_double_pendulum_sys.py", "MyUADoublePendulum", 4, 1, PHYS),
}


def _load_reference_class(monkeypatch, fname, cls):
    for name in [k for k in sys.modules if k == "jax" or k.startswith("jax.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "oracle", "jaxshim"))       # only for its no-op matplotlib
    monkeypatch.syspath_prepend(os.path.join(PKG, "jax_alias"))                # `import jax.numpy` -> class_files.symbolic
    import class_files.systems  # noqa: F401  (this package's: the file's `from .system_base import System` binds here)
    spec = importlib.util.spec_from_file_location("class_files.systems._reference_" + cls, os.path.join(REF, fname))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "class_files.systems"
    spec.loader.exec_module(mod)
    import jax.numpy as jnp
    from class_files import symbolic
    assert jnp is symbolic
    return getattr(mod, cls)


@pytest.mark.parametrize("kind", sorted(CASES))
def test_reference_system_file_loads_unchanged_and_compiles(kind, oracle, monkeypatch, tmp_path):
    import sympy as sp
    from class_files import codegen
    from class_files.systems.system_base import System
    fname, cls_name, n, m, phys = CASES[kind]
    cls = _load_reference_class(monkeypatch, fname, cls_name)
    assert issubclass(cls, System) and cls.__module__.startswith("class_files.systems._reference_")
    rng = np.random.default_rng(3)
    Q, R, Qf = np.diag(rng.uniform(0.5, 2, n)), np.diag(rng.uniform(0.5, 2, m)), np.diag(rng.uniform(5, 20, n))
    xt = rng.uniform(-1, 1, n)
    s = cls(dt=0.01, x_target=xt, Q=Q, R=R, Q_f=Qf, integrator="rk4", **phys)     # the reference's own constructor
    assert s._is_user_defined() and s._device_model()[0] == "user"
    # trace -> analytic derivatives -> device model -> cubin, all in this process
    monkeypatch.setattr(codegen, "CACHE", str(tmp_path))
    text, gn, gm = codegen.generate_header(s)
    assert (gn, gm) == (n, m) and f"from {cls_name}._f_cont_fcn" in text
    cubin, names, _, _ = codegen.compile_module(s)
    assert cubin[:4] == b"\x7fELF" and len(names) == 6
    # the traced expressions against the oracle's restatement of the same file
    _, _, xs, us, f, l, lf = codegen._trace(s)
    fn = sp.lambdify([xs, us], f, "math")
    An = sp.lambdify([xs, us], [[sp.diff(fi, v) for v in xs] for fi in f], "math")
    Bn = sp.lambdify([xs, us], [[sp.diff(fi, v) for v in us] for fi in f], "math")
    ln, lfn = sp.lambdify([xs, us], l, "math"), sp.lambdify([xs], lf, "math")
    p = oracle.make_problem(kind, "rk4", 10, 0.01, Q, R, Qf, xt, phys)
    for _ in range(20):
        x, u = rng.uniform(-2, 2, n), rng.uniform(-2, 2, m)
        Ao, Bo = oracle.f_cont_jac(p, x, u)
        assert np.allclose(np.array(An(x, u), dtype=float), Ao, rtol=1e-11, atol=1e-12)
        assert np.allclose(np.array(Bn(x, u), dtype=float), Bo, rtol=1e-11, atol=1e-12)
        xd = np.empty(n)
        import ctypes as C
        oracle.lib().orc_f_cont(C.byref(p), 0, 0.0, x.ctypes.data_as(C.POINTER(C.c_double)),
                                u.ctypes.data_as(C.POINTER(C.c_double)), xd.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.allclose(np.array(fn(x, u), dtype=float), xd, rtol=1e-12, atol=1e-13)
        assert abs(ln(x, u) - oracle.l(p, x, u)) <= 1e-12 * abs(oracle.l(p, x, u))
        assert abs(lfn(x) - oracle.lf(p, x)) <= 1e-12 * abs(oracle.lf(p, x))
