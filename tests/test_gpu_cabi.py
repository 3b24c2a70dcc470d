"""The C ABI bound INDEPENDENTLY of class_files/_cabi.py (examples/cabi_binding.py: cffi over include/ilqr_b200.h, the
INTEGRATION.md stub with torch buffers): entry points the Python classes do not use -- ilqr_backward on the documented
[N][n][n][B] / [N][n][m][B] linearization, ilqr_forward_linesearch -- against the oracle, and two handles on two
streams at once."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from helpers import UA_OL, cfg2_x0, ua_oracle_problem, ua_system

sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_header_parses_and_every_symbol_is_exported():
    import cabi_binding as cb
    ffi, lib = cb.load()
    names = cb.declared_functions()
    assert len(names) >= 22 and "ilqr_solve" in names and "ilqr_fp64_peak" in names
    for name in names:
        assert getattr(lib, name) is not None, name
    assert lib.ILQR_NMAX == 12 and lib.ILQR_MMAX == 4 and lib.ILQR_UA_DOUBLE_PENDULUM == 2 and lib.ILQR_E_WORKSPACE == -3
    assert ffi.sizeof("ilqr_problem_t") == 10 * 4 + 8 * (4 + 16 + 144 + 16 + 144 + 12 + 144 + 144 + 48 + 1 + 4)


def _ua_solver(cb, ffi, lib, N, B, **kw):
    p = UA_OL
    ph = p["phys"]
    phys = [ph[k] for k in ("g", "m1", "m2", "l1", "l2", "d1", "d2", "theta1", "theta2")]
    return cb.Solver(ffi, lib, lib.ILQR_UA_DOUBLE_PENDULUM, lib.ILQR_RK4, 4, 1, N, B, p["dt"], np.diag(p["Q"]).ravel(),
                     np.diag(p["R"]).ravel(), np.diag(p["Q_f"]).ravel(), p["x_target"], phys, **kw)


@pytest.mark.gpu
def test_linearize_backward_and_linesearch_through_the_header(oracle):
    import torch
    import cabi_binding as cb
    ffi, lib = cb.load()
    B, N, n, m = 37, 60, 4, 1                  # ragged batch
    s = _ua_solver(cb, ffi, lib, N, B)
    rng = np.random.default_rng(3)
    x0 = cfg2_x0(B, seed=21)
    U0 = 0.3 * rng.standard_normal((B, m, N))
    p = ua_oracle_problem(oracle, N)
    dev = dict(dtype=torch.float64, device="cuda")
    st = ffi.cast("void *", torch.cuda.current_stream().cuda_stream)
    # nominal by an alpha = 0 rollout with zero gains (ilqr_rollout), batch-innermost layouts of the header
    x0d = torch.as_tensor(x0.T.copy()).cuda()                                     # [n][B]
    Ud = torch.as_tensor(np.ascontiguousarray(U0.transpose(2, 1, 0))).cuda()      # [N][m][B]
    X = torch.zeros((N + 1, n, B), **dev)
    Un, K0, k0 = torch.empty((N, m, B), **dev), torch.zeros((N, m, n, B), **dev), torch.zeros((N, m, B), **dev)
    Xn, cost = torch.empty_like(X), torch.empty((B,), **dev)
    s.check(lib.ilqr_rollout(s.h, ffi.NULL, s.ptr(x0d), 0.0, s.ptr(X), s.ptr(Ud), s.ptr(k0), s.ptr(K0), s.ptr(Xn), s.ptr(Un),
                             s.ptr(cost), st))
    # linearization in the DOCUMENTED layout, then the reverse scan on it
    A, Bd = torch.empty((N, n, n, B), **dev), torch.empty((N, n, m, B), **dev)
    K, k = torch.empty((N, m, n, B), **dev), torch.empty((N, m, B), **dev)
    s.check(lib.ilqr_linearize(s.h, ffi.NULL, s.ptr(Xn), s.ptr(Un), s.ptr(A), s.ptr(Bd), st))
    s.check(lib.ilqr_backward(s.h, s.ptr(Xn), s.ptr(Un), s.ptr(A), s.ptr(Bd), s.ptr(K), s.ptr(k), st))
    # line search: every step size concurrently
    na = 10
    Xc, Uc = torch.empty((na, N + 1, n, B), **dev), torch.empty((na, N, m, B), **dev)
    ca, win = torch.empty((na, B), **dev), torch.empty((B,), dtype=torch.int32, device="cuda")
    s.check(lib.ilqr_forward_linesearch(s.h, ffi.NULL, s.ptr(x0d), s.ptr(Xn), s.ptr(Un), s.ptr(k), s.ptr(K), s.ptr(cost),
                                        s.ptr(Xc), s.ptr(Uc), s.ptr(ca), ffi.cast("int32_t *", win.data_ptr()), st))
    torch.cuda.synchronize()
    Xh, Uh = Xn.permute(2, 1, 0).cpu().numpy(), Un.permute(2, 1, 0).cpu().numpy()
    Ah, Bh = A.permute(3, 0, 1, 2).cpu().numpy(), Bd.permute(3, 0, 1, 2).cpu().numpy()
    Kh, kh = K.permute(3, 0, 1, 2).cpu().numpy(), k.permute(2, 1, 0).cpu().numpy()
    Xch, cah, winh = Xc.permute(3, 0, 2, 1).cpu().numpy(), ca.t().cpu().numpy(), win.cpu().numpy()
    for b in range(B):
        Xo, Uo, c0 = oracle.forward_pass(p, x0[b], 0.0, np.zeros((n, N + 1)), U0[b], np.zeros((m, N)), np.zeros((N, m, n)))
        assert np.max(np.abs(Xh[b] - Xo)) <= 1e-12 * np.max(np.abs(Xo)) and abs(cost[b].item() - c0) <= 1e-12 * c0
        for t in (0, N // 2, N - 1):
            Ao, Bo = oracle.f_jac(p, Xo[:, t], Uo[:, t])
            assert np.max(np.abs(Ah[b, t] - Ao)) <= 1e-12 * np.max(np.abs(Ao))
            assert np.max(np.abs(Bh[b, t] - Bo)) <= 1e-12 * max(np.max(np.abs(Bo)), 1e-3)
        U_ff, Ko = oracle.backward_pass(p, Xo, Uo)
        assert np.max(np.abs(Kh[b] - Ko)) <= 1e-9 * np.max(np.abs(Ko)), b
        assert np.max(np.abs(kh[b] - U_ff)) <= 1e-9 * max(np.max(np.abs(U_ff)), 1e-6), b
        first = -1
        for a in range(na):
            Xa, Ua, c = oracle.forward_pass(p, x0[b], 0.5 ** a, Xo, Uo, U_ff, Ko)
            assert abs(cah[b, a] - c) <= 1e-8 * abs(c), (b, a)          # (gains differ at rounding level; one pass)
            if first < 0 and cah[b, a] <= cost[b].item():
                first = a
            if a == 3:
                assert np.max(np.abs(Xch[b, a] - Xa)) <= 1e-8 * np.max(np.abs(Xa))
        assert winh[b] == first
    s.close()


@pytest.mark.gpu
def test_two_handles_on_two_streams_share_no_state(oracle):
    """the header's contract: one handle per stream, handles share no state.  Two solvers of different size on two
    streams, enqueued alternately, give what each gives alone."""
    import torch
    from class_files.iLQR_class import iLQR
    N = 60
    xa, xb = cfg2_x0(700, seed=31), cfg2_x0(300, seed=32)

    def make(x0, maxiter):
        return iLQR(ua_system(), N * 0.01, torch.as_tensor(x0).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"),
                    maxiter=maxiter, verbose=False)

    alone = []
    for x0, mi in ((xa, 6), (xb, 9)):
        sol = make(x0, mi)
        X, U, cost = sol.optimize_trajectory()
        alone.append([t.clone() for t in (X, U, cost, sol.K, sol.iterations)])
    sa, sb = make(xa, 6), make(xb, 9)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):                       # re-solve a few times while the other stream is busy
        for sol, st in ((sa, s1), (sb, s2)):
            with torch.cuda.stream(st):
                sol.reset_state()
                sol._U.zero_()
                sol.solve_device(sync=False)
    torch.cuda.synchronize()
    for sol, ref in ((sa, alone[0]), (sb, alone[1])):
        got = (sol.X, sol.U, sol.cost, sol.K, sol.iterations)
        for g, r in zip(got, ref):
            assert torch.equal(g, r)
