"""Exploratory per-kernel timings on one GPU (not the bench contract; see bench.py)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files.iLQR_class import iLQR   # noqa: E402
from class_files import _device as D      # noqa: E402
from helpers import ua_system, cfg2_x0    # noqa: E402


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    integ = sys.argv[3] if len(sys.argv) > 3 else "rk4"
    iters = int(os.environ.get("QG_ITERS", "10"))
    n_alpha = int(os.environ.get("QG_ALPHAS", "10"))
    if integ == "ltv":
        from class_files.systems.ltv_sys import MyLTVSystem
        s = MyLTVSystem.synthetic()
        rng = np.random.default_rng(102)
        x0 = rng.standard_normal((B, 12))
        phi = rng.uniform(0, 2 * np.pi, B)
        n, m = 12, 4
        sol = iLQR(s, N * 0.01, x0, np.zeros((m, N)), tol=0.0, maxiter=iters, verbose=False, phi=phi, n_alpha=n_alpha)
        bytes_k = {"linearize": 1792, "backward": 2080, "rollout": 672}
    else:
        n, m = 4, 1
        x0 = cfg2_x0(B)
        if os.environ.get("QG_SHARD"):          # "r,w": rank r's shard of the w*B batch bench.py uses at --gpus w
            r, w = (int(v) for v in os.environ["QG_SHARD"].split(","))
            x0 = cfg2_x0(B * w)[r * B:(r + 1) * B]
        sol = iLQR(ua_system(integ), N * 0.01, x0, np.zeros((1, N)), tol=0.0, maxiter=iters, verbose=False, n_alpha=n_alpha)
        bytes_k = {"linearize": 240, "backward": 240, "rollout": 120}
    phi_p = D.ptr(sol._phi)
    h = sol._handle
    lib, ws = h.lib, h.workspace()
    print(f"B={B} N={N} {integ} workspace {ws.numel()/1e6:.1f} MB")
    st = D.stream_ptr()
    # one solve to get a representative nominal
    t0 = time.time(); tot = sol.solve_device(); t1 = time.time()
    print(f"first solve: {tot} traj-iters in {t1-t0:.3f}s  status {np.bincount(sol.status, minlength=3)}")
    if integ == "ltv" or os.environ.get("QG_SOLVE_ONLY"):
        return profiled_solves(sol, iters)
    A = torch.empty((N, n, n, B), dtype=torch.float64, device="cuda")
    Bd = torch.empty((N, n, m, B), dtype=torch.float64, device="cuda")
    Xc = torch.empty((n_alpha, N + 1, n, B), dtype=torch.float64, device="cuda")
    Uc = torch.empty((n_alpha, N, m, B), dtype=torch.float64, device="cuda")
    ca = torch.empty((n_alpha, B), dtype=torch.float64, device="cuda")
    win = torch.empty((B,), dtype=torch.int32, device="cuda")
    K2, k2 = torch.empty_like(sol._K), torch.empty_like(sol._k)
    p = D.ptr
    f_lin = lambda: lib.ilqr_linearize(h.h, phi_p, p(sol._X), p(sol._U), p(A), p(Bd), st)
    f_bwd = lambda: lib.ilqr_backward(h.h, p(sol._X), p(sol._U), p(A), p(Bd), p(K2), p(k2), st)
    f_fwd = lambda: lib.ilqr_forward_linesearch(h.h, phi_p, p(sol._x0), p(sol._X), p(sol._U), p(sol._k), p(sol._K),
                                                p(sol._cost), p(Xc), p(Uc), p(ca), p(win), st)
    f_one = lambda: lib.ilqr_rollout(h.h, phi_p, p(sol._x0), 1.0, p(sol._X), p(sol._U), p(sol._k), p(sol._K),
                                     p(Xc), p(Uc), p(ca), st)
    for name, f, bytes_per in (("linearize", f_lin, bytes_k["linearize"]), ("backward", f_bwd, bytes_k["backward"]),
                               (f"rollout x{n_alpha}", f_fwd, bytes_k["rollout"]), ("rollout x1", f_one, bytes_k["rollout"])):
        best, med = timeit(f)
        print(f"{name:12s} best {best:8.3f} ms  median {med:8.3f} ms   algorithmic {bytes_per*N*B/best/1e6:8.1f} GB/s")

    profiled_solves(sol, iters)


def profiled_solves(sol, iters):
    if os.environ.get("QG_NO_REPS"):
        return
    for rep in range(4):
        if rep == 3:
            sol.set_profiling(True)
        sol.reset_state(); sol._U.zero_()
        torch.cuda.synchronize(); t0 = time.time(); tot = sol.solve_device(); t1 = time.time()
        print(f"solve maxiter={iters}: {tot} traj-iters in {(t1-t0)*1e3:.2f} ms -> {tot/(t1-t0)/1e6:.3f} M traj-iter/s")
    kt = sol.kernel_times()
    nit = max(1, kt["linearize"][1], kt["backward"][1])
    print("  per iteration ms: " + ", ".join(f"{k} {v[0]/nit:.3f}" for k, v in kt.items()) + f"  ({nit} iterations)")


if __name__ == "__main__":
    main()
