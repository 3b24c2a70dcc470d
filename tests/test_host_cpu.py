"""CPU-only checks: the C-ABI library loads and exports every symbol include/ilqr_b200.h declares (no
compute calls without a GPU), host-side argument checking mirrors the reference's, the horizon rule,
and the struct layout the ctypes binding assumes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, PKG


def header_functions():
    src = open(os.path.join(ROOT, "include", "ilqr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_binding_expects():
    from class_files import _cabi
    assert header_functions() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from class_files import _cabi
    assert os.path.exists(_cabi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = _cabi.load()
    for name in header_functions():
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ilqr_version()
    assert lib.ilqr_strerror(-3).decode().startswith("workspace")


def test_problem_struct_matches_header_layout(tmp_path):
    """sizeof/offsetof of ilqr_problem_t as gcc sees include/ilqr_b200.h == the ctypes mirror"""
    import subprocess
    from class_files import _cabi
    fields = [f[0] for f in _cabi.Problem._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ilqr_b200.h"\nint main(void){\n'
                   'printf("%zu\\n", sizeof(ilqr_problem_t));\n' +
                   "".join(f'printf("%zu\\n", offsetof(ilqr_problem_t, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == C.sizeof(_cabi.Problem)
    for f, off in zip(fields, vals[1:]):
        assert getattr(_cabi.Problem, f).offset == off, f


def test_create_rejects_bad_problems_without_touching_cuda():
    """ilqr_create validates before any CUDA call, so the ValueError counterparts work on a CPU box."""
    from class_files import _cabi
    lib = _cabi.load()
    p = _cabi.Problem()
    h = C.c_void_p()
    p.model, p.integrator, p.n, p.m, p.N, p.B, p.n_alpha, p.maxiter, p.dt = 2, 7, 4, 1, 10, 1, 10, 5, 0.01
    assert lib.ilqr_create(C.byref(p), C.byref(h)) == -1            # unknown integrator
    p.integrator, p.n = 2, 3
    assert lib.ilqr_create(C.byref(p), C.byref(h)) == -1            # dimension/model mismatch
    p.n, p.n_alpha = 4, 99
    assert lib.ilqr_create(C.byref(p), C.byref(h)) == -1
    assert lib.ilqr_destroy(None) == -1 and lib.ilqr_workspace_bytes(None) == 0


def test_unknown_integrator_and_shapes_raise_like_the_reference():
    from class_files.systems.pendulum_sys import MyPendulum
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum
    from class_files.iLQR_class import iLQR
    kw = dict(dt=0.01, x_target=np.array([np.pi, 0.0]), Q=np.eye(2), R=np.eye(1), Q_f=np.eye(2))
    with pytest.raises(ValueError, match="Unknown integrator: 'rk5'"):          # system_base.py:197-198
        MyPendulum(integrator="rk5", **kw)
    s = MyPendulum(integrator="backward_euler", **kw)
    assert (s.n_x, s.n_u, s.dt) == (2, 1, 0.01)
    with pytest.raises(ValueError, match=r"U_init must have shape \(1, 400\), but got \(1, 399\)"):   # iLQR_class.py:50-52
        iLQR(s, 4.0, np.array([1.0, 0.0]), np.zeros((1, 399)))
    with pytest.raises(ValueError, match="x_0 must have shape"):
        iLQR(s, 4.0, np.zeros(3), np.zeros((1, 400)))
    ua = MyUADoublePendulum(dt=0.01, x_target=np.zeros(4), Q=np.eye(4), R=np.eye(1), Q_f=np.eye(4))
    with pytest.raises(ValueError, match="Q must have shape"):
        MyUADoublePendulum(dt=0.01, x_target=np.zeros(4), Q=np.eye(3), R=np.eye(1), Q_f=np.eye(4)).make_problem(10, 1)
    p = ua.make_problem(N=500, B=7, maxiter=3)
    assert (p.model, p.integrator, p.n, p.m, p.N, p.B, p.maxiter) == (2, 2, 4, 1, 500, 7, 3)
    assert list(p.phys)[:9] == [9.81, 1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0, 0.0]


def test_no_cpu_fallback():
    """Without a CUDA device the solver refuses to construct (after the reference's own argument checks)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from class_files.systems.pendulum_sys import MyPendulum
    from class_files.iLQR_class import iLQR
    s = MyPendulum(dt=0.01, x_target=np.array([np.pi, 0.0]), Q=np.eye(2), R=np.eye(1), Q_f=np.eye(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        iLQR(s, 1.0, np.zeros(2), np.zeros((1, 100)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.f_fcn(np.zeros(2), np.zeros(1))


@pytest.mark.parametrize("T,N", [(1.0, 100), (2.0, 200), (3.0, 300), (4.0, 400), (5.0, 500), (8.0, 800), (0.5, 50)])
def test_horizon_rule(T, N, oracle):
    """N = len(arange(0, T+dt, dt)) - 1 (iLQR_class.py:46-47)."""
    assert oracle.horizon(T, 0.01) == N
    assert len(np.arange(0, T + 0.01, 0.01)) - 1 == N


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ilqr_oracle" not in txt and "jaxshim" not in txt, os.path.join(dirpath, f)


def test_trig_table_header_is_reproducible_and_accurate(tmp_path):
    """csrc/ilqr_trig_table.cuh (the table and constants of sincos_tab, the sincos of the large-batch FP64 kernels) is what
    scripts/gen_trig_table.py writes, and the algorithm it parameterises -- restated here in numpy without FMA -- is
    within 2.5e-16 of sin and cos for arguments up to 1e3 (the device code, with FMA, measures 1.1e-16)."""
    import subprocess
    import sys
    mp = pytest.importorskip("mpmath")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "iterative-linear-quadratic-regulator_b200", "csrc", "ilqr_trig_table.cuh")
    committed = open(hdr).read()
    again = str(tmp_path / "ilqr_trig_table.cuh")
    subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_trig_table.py"), again], check=True, capture_output=True)
    assert open(again).read() == committed
    const = {k: float(v) for k, v in re.findall(r"#define ILQR_TRIG_(\w+) (\S+)", committed)}
    tab = np.array([[float(a), float(b)] for a, b in re.findall(r"\{(\S+), (\S+)\},", committed)])
    n = int(const["N"])
    assert tab.shape == (n, 2) and n == 512
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-40, 40, 20000), rng.uniform(-1e3, 1e3, 5000)])
    k = np.rint(x * const["INV_H"])
    r = (x - k * const["H_HI"]) - k * const["H_LO"]
    z = r * r
    sr = r + (r * z) * (const["S1"] + z * const["S2"])
    cm = z * (const["C1"] + z * const["C2"])
    S, C = tab[k.astype(np.int64) & (n - 1)].T
    s, c = S + (S * cm + C * sr), C + (C * cm - S * sr)
    mp.mp.dps = 40
    es = max(abs(mp.mpf(float(a)) - mp.sin(mp.mpf(float(b)))) for a, b in zip(s, x))
    ec = max(abs(mp.mpf(float(a)) - mp.cos(mp.mpf(float(b)))) for a, b in zip(c, x))
    assert es < 2.5e-16 and ec < 2.5e-16, (float(es), float(ec))


def test_sass_of_the_shipped_library_has_what_the_design_claims():
    """the thread-per-trajectory Riccati kernel fills its ring with the TMA unit's linear copies
    (UBLKCP) completing on mbarriers (SYNCS ... TRANS64), the fused kernel hands its ring over through mbarriers, the LTV
    Riccati kernel runs on the FP64 tensor cores (DMMA).  Read from the sm_100a cubin inside libilqr_b200.so."""
    import shutil
    import subprocess
    from class_files import _cabi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    functions = re.split(r"\n\s*Function : ", sass)
    by = lambda frag: [f for f in functions if frag in f.split("\n", 1)[0]]
    k2 = by("_ZN4ilqr15backward_kernel")                              # (mangled: not fused_backward_kernel)
    assert k2 and all("UBLKCP" in f and "SYNCS.ARRIVE.TRANS64" in f and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in f for f in k2)
    assert all("LDGSTS" in f for f in k2)                              # the per-thread ring is still there (small batches)
    fused = by("fused_backward")
    assert fused and all("SYNCS.PHASECHK.TRANS64.TRYWAIT" in f for f in fused)
    mma = by("backward_ltv_mma_kernel")
    assert mma and all("DMMA.8x8x4" in f for f in mma)
