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
    iters = 10
    x0 = cfg2_x0(B)
    sol = iLQR(ua_system(integ), N * 0.01, x0, np.zeros((1, N)), tol=0.0, maxiter=iters, verbose=False)
    h = sol._handle
    lib, ws = h.lib, h.workspace()
    print(f"B={B} N={N} {integ} workspace {ws.numel()/1e6:.1f} MB")
    st = D.stream_ptr()
    # one solve to get a representative nominal
    t0 = time.time(); tot = sol.solve_device(); t1 = time.time()
    print(f"first solve: {tot} traj-iters in {t1-t0:.3f}s  status {np.bincount(sol.status, minlength=3)}")
    n, m = 4, 1
    A = torch.empty((N, n, n, B), dtype=torch.float64, device="cuda")
    Bd = torch.empty((N, n, m, B), dtype=torch.float64, device="cuda")
    Xc = torch.empty((10, N + 1, n, B), dtype=torch.float64, device="cuda")
    Uc = torch.empty((10, N, m, B), dtype=torch.float64, device="cuda")
    ca = torch.empty((10, B), dtype=torch.float64, device="cuda")
    win = torch.empty((B,), dtype=torch.int32, device="cuda")
    K2, k2 = torch.empty_like(sol._K), torch.empty_like(sol._k)
    p = D.ptr
    f_lin = lambda: lib.ilqr_linearize(h.h, None, p(sol._X), p(sol._U), p(A), p(Bd), st)
    f_bwd = lambda: lib.ilqr_backward(h.h, p(sol._X), p(sol._U), p(A), p(Bd), p(K2), p(k2), st)
    f_fwd = lambda: lib.ilqr_forward_linesearch(h.h, None, p(sol._x0), p(sol._X), p(sol._U), p(sol._k), p(sol._K),
                                                p(sol._cost), p(Xc), p(Uc), p(ca), p(win), st)
    f_one = lambda: lib.ilqr_rollout(h.h, None, p(sol._x0), 1.0, p(sol._X), p(sol._U), p(sol._k), p(sol._K),
                                     p(Xc), p(Uc), p(ca), st)
    for name, f, bytes_per in (("linearize", f_lin, 240), ("backward", f_bwd, 240), ("rollout x10", f_fwd, 120),
                               ("rollout x1", f_one, 120)):
        best, med = timeit(f)
        print(f"{name:12s} best {best:8.3f} ms  median {med:8.3f} ms   algorithmic {bytes_per*N*B/best/1e6:8.1f} GB/s")

    def full():
        sol.X = np.zeros((4, N + 1)); sol.K = np.zeros((N, 1, 4)); sol.U_ff = np.zeros((1, N)); sol.U = np.zeros((1, N))
        return sol.solve_device()
    full(); torch.cuda.synchronize()
    for _ in range(3):
        sol.X = np.zeros((4, N + 1)); sol.K = np.zeros((N, 1, 4)); sol.U_ff = np.zeros((1, N)); sol.U = np.zeros((1, N))
        torch.cuda.synchronize(); t0 = time.time(); tot = sol.solve_device(); t1 = time.time()
        print(f"solve maxiter={iters}: {tot} traj-iters in {(t1-t0)*1e3:.2f} ms -> {tot/(t1-t0)/1e6:.3f} M traj-iter/s")


if __name__ == "__main__":
    main()
