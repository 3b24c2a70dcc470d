"""Exploration: time one batched solve for every shipped system / integrator / precision (A/B two builds with
ILQR_B200_LIB=...).   python scripts/sweep_systems.py [B] [N]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files.iLQR_class import iLQR                                      # noqa: E402
from class_files.systems.pendulum_sys import MyPendulum                      # noqa: E402
from class_files.systems.double_pendulum_sys import MyDoublePendulum         # noqa: E402
from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum    # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
phys = dict(g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1, theta1=1 / 12, theta2=1 / 12)
rng = np.random.default_rng(5)
cases = []
for integ in ("euler", "midpoint", "rk4", "backward_euler"):
    for dtype in ("float64", "float32"):
        cases.append(("pendulum", integ, dtype))
        cases.append(("double", integ, dtype))
        cases.append(("ua", integ, dtype))
for kind, integ, dtype in cases:
    if kind == "pendulum":
        s = MyPendulum(dt=0.01, x_target=np.array([np.pi, 0.0]), Q=np.diag([1.0, 1.0]), R=np.diag([1.0]), Q_f=np.diag([100.0, 10.0]),
                       d=0.1, integrator=integ, dtype=dtype)
    else:
        cls = MyDoublePendulum if kind == "double" else MyUADoublePendulum
        m = 2 if kind == "double" else 1
        s = cls(dt=0.01, x_target=np.array([np.pi, 0.0, 0.0, 0.0]), Q=np.diag([1.0, 1.0, 0.1, 0.1]), R=np.diag([1.0] * m),
                Q_f=np.diag([1000.0, 1000.0, 100.0, 100.0]), integrator=integ, dtype=dtype, **phys)
    x0 = rng.uniform(-1.0, 1.0, (B, s.n_x))
    sol = iLQR(s, N * 0.01, x0, np.zeros((s.n_u, N)), tol=0.0, maxiter=5, verbose=False)
    ts = []
    for rep in range(4):
        sol.reset_state(); sol._U.zero_()
        torch.cuda.synchronize(); t0 = time.time(); tot = sol.solve_device(); torch.cuda.synchronize(); ts.append(time.time() - t0)
    print(f"{kind:9s} {integ:15s} {dtype:8s} {min(ts[1:]) * 1e3:8.3f} ms  ({tot} traj-iters)", flush=True)
