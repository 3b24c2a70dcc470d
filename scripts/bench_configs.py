#!/usr/bin/env python
"""Secondary measurements on one B200 (not the bench.py contract): BASELINE.json configs 3, 4, 5 at full size and
the FP32 mode of config 2.  Prints one JSON object; CUDA-event timed, inputs resident in HBM.

    python scripts/bench_configs.py > profiles/<round>_configs.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files.iLQR_class import iLQR                      # noqa: E402
from class_files.mpc import run_mpc                          # noqa: E402
from class_files.chunked import solve_chunked                # noqa: E402
from class_files.systems.ltv_sys import MyLTVSystem          # noqa: E402
from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum   # noqa: E402
from helpers import ua_system, cfg2_x0                       # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = [fn() for _ in range(reps)]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out[-1]


def cfg2(B, dtype, iters=10):
    tdt = torch.float64 if dtype == "float64" else torch.float32
    x0 = torch.as_tensor(cfg2_x0(B)).to(device="cuda", dtype=tdt)
    sol = iLQR(ua_system(dtype=dtype), 5.0, x0, torch.zeros((1, 500), dtype=tdt, device="cuda"), tol=0.0, maxiter=iters,
               verbose=False)

    def step():
        sol.reset_state(); sol._U.zero_()
        return sol.solve_device(sync=True)
    ms, units = timed(step)
    return {"batch": B, "dtype": dtype, "ms_per_solve": ms, "traj_iter_per_s": units / ms * 1e3}


def cfg3(B=65536, ticks=20):
    phys = dict(g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1, theta1=1 / 12, theta2=1 / 12)
    mk = lambda integ: MyUADoublePendulum(dt=0.01, x_target=np.array([np.pi, 0, 0, 0]), Q=np.diag([5, 5, .1, .1]),
                                          R=np.diag([50.0]), Q_f=np.diag([1000, 1000, 10, 10.0]), integrator=integ, **phys)
    rng = np.random.default_rng(1)
    x0 = torch.as_tensor(rng.standard_normal((B, 4)) * np.array([0.1, 0.1, 0.5, 0.5])).cuda()
    sol = iLQR(mk("rk4"), 2.0, x0, torch.zeros((1, 200), dtype=torch.float64, device="cuda"), maxiter=50, verbose=False, n_alpha=8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = run_mpc(sol, mk("backward_euler"), x0, ticks)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    its = int(r["iterations"].sum().item())
    return {"instances": B, "horizon": 200, "ticks": ticks, "alphas": 8, "seconds": dt, "traj_iterations": its,
            "traj_iter_per_s": its / dt, "mean_iterations_per_tick": its / (B * ticks), "ms_per_tick": dt / ticks * 1e3}


def cfg4(B=262144, N=1000, chunk=32768):
    s = MyLTVSystem.synthetic(seed=2)
    rng = np.random.default_rng(2)
    x0 = torch.as_tensor(rng.standard_normal((B, 12))).cuda()
    phi = torch.as_tensor(rng.uniform(0, 2 * np.pi, B)).cuda()
    U0 = torch.zeros((4, N), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = solve_chunked(s, N * s.dt, x0, U0, chunk, phi=phi, maxiter=2, n_alpha=4)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return {"batch": B, "horizon": N, "chunk": chunk, "iterations": 2, "seconds": dt,
            "traj_iter_per_s": r["total_iterations"] / dt}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
        print(json.dumps(cfg3(ticks=int(sys.argv[2]) if len(sys.argv) > 2 else 20)))
        return
    out = {"gpu": torch.cuda.get_device_name(0)}
    out["cfg2_f64_B4096"] = cfg2(4096, "float64")
    out["cfg2_f32_B4096"] = cfg2(4096, "float32")
    out["cfg5_shard_f64_B131072"] = cfg2(131072, "float64")
    torch.cuda.empty_cache()
    out["cfg5_shard_f32_B131072"] = cfg2(131072, "float32")
    torch.cuda.empty_cache()
    out["cfg3_mpc"] = cfg3()
    torch.cuda.empty_cache()
    out["cfg4_ltv"] = cfg4()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
