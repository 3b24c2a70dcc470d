"""Under-actuated double-pendulum MPC: the loop of the reference's run_iLQR_UA_MPC.py:146-174 written out as there,
then the same closed loop for a batch of instances with every tick on the GPU (class_files.mpc.run_mpc)."""
import time

import numpy as np

import _path  # noqa: F401
from class_files.iLQR_class import iLQR
from class_files.mpc import run_mpc
from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum


def make(integrator):                                          # run_iLQR_UA_MPC.py:17-95
    return MyUADoublePendulum(dt=0.01, x_target=np.array([np.pi, 0.0, 0.0, 0.0]), Q=np.diag([5.0, 5.0, 0.1, 0.1]),
                              R=np.diag([50.0]), Q_f=np.diag([1000.0, 1000.0, 10.0, 10.0]), g=9.81, m1=1.0, m2=1.0, l1=1.0,
                              l2=1.0, d1=0.1, d2=0.1, theta1=1.0 / 12, theta2=1.0 / 12, integrator=integrator)


def main(N_sim=25, B=4096):
    opt, plant = make("rk4"), make("backward_euler")
    T_horizon, N = 2.0, 200
    x0 = np.array([0.05, -0.08, 0.4, -0.3])
    solver = iLQR(opt, T_horizon, x0, np.zeros((1, N)), tol=1e-5, maxiter=50, verbose=False)
    current_x, U_guess = x0, np.zeros((1, N))
    t0 = time.time()
    for k in range(N_sim):                                     # :146
        solver.x_0 = current_x                                 # :148
        solver.U = U_guess                                     # :151
        X_bar, U_bar, cost = solver.optimize_trajectory()      # :154
        uk = U_bar[:, 0]                                       # :157
        current_x = plant.f_fcn(current_x, uk)                 # :161
        U_guess = np.concatenate([U_bar[:, 1:], U_bar[:, -1:]], axis=1)   # :168
    print(f"script-style loop, 1 instance, {N_sim} ticks: {time.time() - t0:.2f} s, state {current_x}")

    rng = np.random.default_rng(1)
    x0b = rng.standard_normal((B, 4)) * np.array([0.1, 0.1, 0.5, 0.5])
    solver_b = iLQR(opt, T_horizon, x0b, np.zeros((1, N)), tol=1e-5, maxiter=50, verbose=False, n_alpha=8)
    t0 = time.time()
    r = run_mpc(solver_b, plant, x0b, N_sim)
    dt = time.time() - t0
    print(f"run_mpc, {B} instances, {N_sim} ticks: {dt:.2f} s ({int(r['iterations'].sum())} trajectory-iterations, "
          f"{r['iterations'].sum() / dt / 1e6:.2f} M/s); mean |x - target| after the last tick "
          f"{np.abs(r['X_sim'][:, :, -1] - np.array([np.pi, 0, 0, 0])).mean():.3f}")


if __name__ == "__main__":
    main()
