"""4096 double-pendulum swing-ups at once (parameters of the reference's run_iLQR_OL_UA_Pendulum.py:17-56)."""
import time

import numpy as np

import _path  # noqa: F401
from class_files.iLQR_class import iLQR
from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum


def main(B=4096):
    sysm = MyUADoublePendulum(dt=0.01, x_target=np.array([np.pi, 0.0, 0.0, 0.0]), Q=np.diag([1.0, 1.0, 0.1, 0.1]),
                              R=np.diag([1.0]), Q_f=np.diag([1000.0, 1000.0, 100.0, 100.0]), g=9.81, m1=1.0, m2=1.0,
                              l1=1.0, l2=1.0, d1=0.1, d2=0.1, theta1=1.0 / 12, theta2=1.0 / 12, integrator="rk4")
    rng = np.random.default_rng(0)
    x0 = np.concatenate([rng.uniform(-np.pi, np.pi, (B, 2)), rng.uniform(-2, 2, (B, 2))], axis=1)
    solver = iLQR(sysm, 5.0, x0, np.zeros((1, 500)), tol=1e-5, maxiter=100, verbose=True,
                  reg_factor=10.0)          # extension: retry failed line searches with a regularised Q_uu
    t0 = time.time()
    X, U, cost = solver.optimize_trajectory()
    dt = time.time() - t0
    print(f"{B} trajectories, {solver.total_iterations} trajectory-iterations in {dt:.2f} s "
          f"({solver.total_iterations / dt / 1e6:.2f} M/s); X {X.shape}, U {U.shape}")


if __name__ == "__main__":
    main()
