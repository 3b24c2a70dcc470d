"""Where the end-to-end step (host numpy in -> host numpy out) spends its time (exploration)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files.iLQR_class import iLQR
from helpers import ua_system, cfg2_x0
B, N = 4096, 500
x0 = cfg2_x0(B); U0 = np.zeros((1, N))
sol = iLQR(ua_system(), 5.0, x0, U0, tol=0.0, maxiter=10, verbose=False)
def lap(name, fn, acc):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.perf_counter() - t; return r
acc = {}
for rep in range(12):
    if rep == 2: acc = {}
    lap("set x0,U", lambda: (setattr(sol, "x_0", x0), setattr(sol, "U", U0)), acc)
    lap("reset", sol.reset_state, acc)
    lap("solve", lambda: sol.solve_device(sync=True), acc)
    X = lap("X", lambda: sol.X, acc)
    U = lap("U", lambda: sol.U, acc)
    c = lap("cost", lambda: sol.cost, acc)
print({k: round(v / 10 * 1e3, 3) for k, v in acc.items()}, "ms")

# bench-style loop: no intermediate syncs, previous results kept alive while the next step runs
def step():
    sol.x_0 = x0; sol.U = U0; sol.reset_state()
    return sol.optimize_trajectory()
for _ in range(3): X, U, c = step()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): X, U, c = step()
torch.cuda.synchronize(); print("bench-style step", (time.perf_counter() - t) / 10 * 1e3, "ms")
t = time.perf_counter()
for _ in range(10):
    X = U = c = None
    X, U, c = step()
torch.cuda.synchronize(); print("results dropped before the next step", (time.perf_counter() - t) / 10 * 1e3, "ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): X, U, c = step()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# a second, device-resident solver alive next to the host one (as in bench.py)
sol2 = iLQR(ua_system(), 5.0, torch.as_tensor(x0).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"), tol=0.0, maxiter=10, verbose=False)
for _ in range(5):
    sol2.reset_state(); sol2._U.zero_(); sol2.solve_device(sync=True)
sol3 = iLQR(ua_system(), 5.0, x0, U0, tol=0.0, maxiter=10, verbose=False)
def step3():
    sol3.x_0 = x0; sol3.U = U0; sol3.reset_state()
    return sol3.optimize_trajectory()
for _ in range(3): X, U, c = step3()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): X, U, c = step3()
torch.cuda.synchronize(); print("fresh host solver after a device solver ran", (time.perf_counter() - t) / 10 * 1e3, "ms")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): X, U, c = step()
torch.cuda.synchronize(); print("first host solver again", (time.perf_counter() - t) / 10 * 1e3, "ms")
