"""Single-pendulum open-loop swing-up, the reference's run_iLQR_open_loop.py:16-105 without the plots."""
import time

import numpy as np

import _path  # noqa: F401
from class_files.iLQR_class import iLQR
from class_files.systems.pendulum_sys import MyPendulum


def main():
    dt, T = 0.01, 4.0                                        # run_iLQR_open_loop.py:16-24
    x_target = np.array([np.pi, 0.0])
    Q, R, Q_f = np.diag([1.0, 1.0]), np.diag([1.0]), np.diag([0.0, 0.0])
    pendulum = MyPendulum(dt=dt, x_target=x_target, Q=Q, R=R, Q_f=Q_f, g=9.81, l=1.0, d=0.0, integrator="backward_euler")
    x_0 = np.array([1.0, 0.0])
    N = int(T / dt)
    solver = iLQR(pendulum, T, x_0, np.zeros((1, N)), tol=1e-5, maxiter=100, verbose=True)
    t0 = time.time()
    X, U, cost = solver.optimize_trajectory()
    X.block_until_ready()                                     # as upstream (:83,93)
    print(f"solved in {time.time() - t0:.3f} s: cost {float(cost):.4f}, final state {X[:, -1]}, target {x_target}")


if __name__ == "__main__":
    main()
