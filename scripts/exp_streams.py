"""Experiment: does running S independent sub-batches on S streams overlap the latency-bound phases?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]
from class_files.iLQR_class import iLQR
from helpers import ua_system, cfg2_x0
B, N, IT = 4096, 500, 10
x0 = cfg2_x0(B)
for S in (1, 2, 4, 8):
    bs = B // S
    sols = [iLQR(ua_system(), 5.0, torch.as_tensor(x0[i*bs:(i+1)*bs]).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"), tol=0.0, maxiter=IT, verbose=False) for i in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    def run():
        for s, st in zip(sols, streams):
            with torch.cuda.stream(st):
                s.reset_state(); s._U.zero_()
                s.solve_device(sync=False)
        torch.cuda.synchronize()
    for _ in range(3): run()
    t0 = time.perf_counter()
    for _ in range(10): run()
    dt = (time.perf_counter() - t0) / 10
    print(f"S={S}: {dt*1e3:.2f} ms per {B*IT} traj-iters -> {B*IT/dt/1e6:.2f} M/s")
