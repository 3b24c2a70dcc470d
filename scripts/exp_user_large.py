"""A user-defined system (the tests' cart-pole with the barrier cost: NVRTC-compiled generic kernels, state-dependent cost
Hessians) at a large batch: per-kernel times with the per-thread and the bulk-copy ring of the Riccati scan."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files import symbolic as jnp                      # noqa: E402
from class_files.iLQR_class import iLQR                      # noqa: E402
from class_files.systems.system_base import System           # noqa: E402
from user_systems import CARTPOLE as p, make_cartpole_class  # noqa: E402

B, N, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 200, 4
s = make_cartpole_class(System, jnp)(dt=p["dt"], x_target=np.array(p["x_target"]), Q=np.diag(p["Q"]), R=np.diag(p["R"]),
                                     Q_f=np.diag(p["Q_f"]), **{k: p[k] for k in ("mc", "mp", "l", "g", "b", "p_max", "w_bar")})
x0 = np.random.default_rng(3).uniform(-0.5, 0.5, (B, 4))
for bulk in ("0", "1"):
    os.environ["ILQR_BACKWARD_BULK"] = bulk
    sol = iLQR(s, N * 0.01, torch.as_tensor(x0).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"), tol=0.0,
               maxiter=iters, verbose=False)
    for rep in range(3):
        if rep == 2:
            sol.set_profiling(True)
        sol.reset_state(); sol._U.zero_()
        torch.cuda.synchronize(); t0 = time.time(); tot = sol.solve_device(); t1 = time.time()
    kt = sol.kernel_times()
    nit = max(1, kt["linearize"][1], kt["backward"][1])
    print(f"B={B} N={N} user cart-pole, ILQR_BACKWARD_BULK={bulk}: {tot/(t1-t0)/1e6:.2f} M traj-iter/s; per iteration ms: "
          + ", ".join(f"{k} {v[0]/nit:.3f}" for k, v in kt.items()), flush=True)
    del sol
