#!/usr/bin/env python
"""bench.py -- batched iLQR trajectory-iterations/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): under-actuated double pendulum swing-up, n=4, m=1, dt=0.01,
T=5 => N=500, rk4, parameters of run_iLQR_OL_UA_Pendulum.py:27-56, B=4096 seeded random initial
states PER GPU (weak scaling: the batch shards by trajectory, no data-path collective; per-shard
costs/flags are all-gathered over NCCL after each step).
One "step" = one batched solve from fresh solver state with maxiter=ITERS, tol=0: the alpha=0
initial rollout plus ITERS x (linearize, backward Riccati, 10-alpha line search, select/convergence).
Units = sum over trajectories of iterations actually executed (device-side counters); the initial
rollout is inside the timed region but is not counted as an iteration.
`value`  : device-resident (x0/U already in HBM), CUDA-event timed, max over ranks.
`e2e`    : the same solve through the reference-style Python API with HOST numpy inputs and outputs
           (H2D of x0/U_init and D2H of X, U, cost inside the timed region).
The --impl reference arm times the CPU oracle port (oracle/ilqr_oracle.c: the reference is pure
Python on JAX, absent from this image and from the GPU box) on all host cores, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200"), os.path.join(ROOT, "tests")]

METRIC = "batched iLQR traj-iterations/sec (double pendulum, N=500)"
UNIT = "traj-iter/s"
N_H, T_H, DT = 500, 5.0, 0.01
ITERS = 10
N_ALPHA = 10
# algorithmic bytes per trajectory-timestep, FP64 (DESIGN.md "Roofline accounting")
BYTES = {"linearize": 240, "backward": 240, "rollout": 120}


def cfg2_x0(count, seed=0):
    rng = np.random.default_rng(seed)
    x0 = np.empty((count, 4))
    x0[:, :2] = rng.uniform(-np.pi, np.pi, size=(count, 2))
    x0[:, 2:] = rng.uniform(-2.0, 2.0, size=(count, 2))
    return x0


UA = dict(Q=[1.0, 1.0, 0.1, 0.1], R=[1.0], Q_f=[1000.0, 1000.0, 100.0, 100.0], x_target=[np.pi, 0.0, 0.0, 0.0],
          phys=dict(g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1, theta1=1.0 / 12, theta2=1.0 / 12))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.25:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ilqr_oracle as O
    threads = O.max_threads()
    sample = args.cpu_sample
    p = O.make_problem("ua", "rk4", N_H, DT, UA["Q"], UA["R"], UA["Q_f"], UA["x_target"], UA["phys"], tol=0.0,
                       maxiter=ITERS, n_alpha=N_ALPHA)
    x0 = cfg2_x0(args.batch)[:sample]
    U0 = np.zeros((sample, 1, N_H))
    for _ in range(max(args.warmup, 1)):
        O.optimize_batch(p, x0[: max(threads, 8)], U0[: max(threads, 8)], nthreads=threads)
    units, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        r = O.optimize_batch(p, x0, U0, nthreads=threads)
        units += int(r["iters"].sum())
    dt = time.perf_counter() - t0
    val = units / dt
    desc = f"{sample} of the {args.batch} cfg2 trajectories x {ITERS} iterations per step, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, world):
    return {"workload": "cfg2: UA double-pendulum open-loop swing-up (run_iLQR_OL_UA_Pendulum.py params), n=4 m=1 "
                        f"N={N_H} rk4, B={args.batch} seeded initial states per GPU",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "horizon": N_H,
            "iterations_per_step": ITERS, "line_search_alphas": N_ALPHA, "tol": 0.0, "sharding": f"batch x{world}",
            "l2_policy": "working set per step (A,B,X,U,K,k + 10 candidate slabs = 1.3 GB) exceeds the 126 MB L2; "
                         "no explicit flush"}


def large_batch_line(torch, iLQR, sysm, BL, steps=3, warmup=2):
    """The bench step at BL trajectories on this GPU (x0 from the same generator): traj-iter/s and per-kernel
    times per iLQR iteration.  At this size the backward pass streams its linearization at HBM speed."""
    x0 = torch.as_tensor(cfg2_x0(BL, seed=1)).cuda()
    U0 = torch.zeros((1, N_H), dtype=torch.float64, device="cuda")
    sol = iLQR(sysm, T_H, x0, U0, tol=0.0, maxiter=ITERS, verbose=False, n_alpha=N_ALPHA)

    def step():
        sol.reset_state()
        sol._U.zero_()
        return sol.solve_device(sync=True)
    for _ in range(warmup):
        step()
    sol.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    units = sum(step() for _ in range(steps))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kt = sol.kernel_times()
    nit = max(1, kt["linearize"][1])
    kern = {}
    for name in ("linearize", "backward", "rollout"):
        avg = kt[name][0] / nit
        alg = BYTES[name] * N_H * BL + (8 * N_ALPHA * BL if name == "rollout" else 0)
        kern[name] = {"ms_per_iteration": avg, "achieved_GBps": alg / avg / 1e6}
    return {"batch": BL, "value": units / (ms * 1e-3), "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
            "line_search": "lazy waves of 2,2,2,4 step sizes over compacted lists", "kernels": kern,
            "roofline_backward": {"bound": "hbm", "kernel": "backward", "unit": "GB/s",
                                  "achieved": kern["backward"]["achieved_GBps"]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="trajectories per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="trajectories per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="secondary runs only (e.g. --batch 131072 on 8 GPUs = config 5): skip the host-in/host-out arm")
    ap.add_argument("--large-batch", type=int, default=131072,
                    help="N=1 only: also time a few steps at this batch (config 5's per-GPU shard at 8 GPUs), where "
                         "the passes are throughput bound; reported under 'large_batch'.  0 disables.")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, in front of the one JSON line this prints
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from class_files.iLQR_class import iLQR
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum

    sysm = MyUADoublePendulum(dt=DT, x_target=np.array(UA["x_target"]), Q=np.diag(UA["Q"]), R=np.diag(UA["R"]),
                              Q_f=np.diag(UA["Q_f"]), integrator="rk4", **UA["phys"])
    B = args.batch
    x0_all = cfg2_x0(B * world)
    x0_host = np.ascontiguousarray(x0_all[rank * B:(rank + 1) * B])        # this rank's contiguous shard
    U_host = np.zeros((1, N_H))
    # ---- device-resident arm: torch tensors in, nothing crosses PCIe in the timed region --------------
    x0_dev = torch.as_tensor(x0_host).cuda()
    U_dev = torch.zeros((1, N_H), dtype=torch.float64, device="cuda")
    sol = iLQR(sysm, T_H, x0_dev, U_dev, tol=0.0, maxiter=ITERS, verbose=False, n_alpha=N_ALPHA)
    # the path's only exchange: per-shard cost and exit status, packed into ONE NCCL all-gather per step
    pack = torch.empty((2, B), dtype=torch.float64, device="cuda") if world > 1 else None
    gathered = torch.empty((world, 2, B), dtype=torch.float64, device="cuda") if world > 1 else None

    def step_device():
        sol.reset_state()
        sol._U.zero_()
        units = sol.solve_device(sync=True)
        if world > 1:
            pack[0].copy_(sol._cost)
            pack[1].copy_(sol._status)
            dist.all_gather_into_tensor(gathered, pack)
        return units

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    l0 = sol.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.time()
    e0.record()
    units = 0
    for _ in range(args.steps):
        units += step_device()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = sol.launches() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    # per-kernel breakdown: the same steps again with CUDA events chained between the launches (kept out of
    # the timed region above: the extra event records cost ~3 % at this batch)
    sol.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(max(1, min(args.steps, 10))):
        step_device()
    pe1.record()
    torch.cuda.synchronize()
    ms_prof = pe0.elapsed_time(pe1)
    ktimes = sol.kernel_times()
    sol.set_profiling(False)

    # ---- end-to-end arm: host numpy in, host numpy out through the public API -----------------------
    te, units_e, h2d, d2h = float("nan"), 0, 0, 0
    sol_h = None
    if not args.skip_e2e:
        sol_h = iLQR(sysm, T_H, x0_host, U_host, tol=0.0, maxiter=ITERS, verbose=False, n_alpha=N_ALPHA)

        def step_e2e():
            sol_h.x_0 = x0_host                   # H2D
            sol_h.U = U_host                      # H2D
            sol_h.reset_state()
            X, U, cost = sol_h.optimize_trajectory()   # D2H of X, U, cost
            return sol_h.total_iterations, X, U, cost

        for _ in range(max(3, args.warmup)):
            # keep the results alive across steps exactly as the timed loop does, so that the pinned staging
            # buffers of the steady state (two per output) exist before the clock starts
            u, X, U, cost = step_e2e()
        barrier()
        te0 = time.perf_counter()
        for _ in range(args.steps):
            u, X, U, cost = step_e2e()
            units_e += u
        barrier()
        te = time.perf_counter() - te0
        h2d = x0_host.nbytes + U_host.nbytes
        d2h = X.nbytes + U.nbytes + cost.nbytes

    # ---- N=1 only: the same solve at config 5's per-GPU shard size (HBM/FP64-throughput-bound regime) ------
    large = None
    if world == 1 and args.large_batch > 0:
        del sol_h
        torch.cuda.empty_cache()
        large = large_batch_line(torch, iLQR, sysm, args.large_batch)

    if world > 1:
        t = torch.tensor([ms, te, float(units), float(units_e)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, te = float(tmax[0]), float(tmax[1])
        units, units_e = int(tsum[2]), int(tsum[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    def kinfo(name):
        tot, cnt = ktimes[name]
        if cnt == 0:
            return None
        # per iLQR iteration (the lazy line-search schedule launches the rollout kernel once per wave)
        avg = tot / max(1, ktimes["linearize"][1])
        alg = BYTES[name] * N_H * B + (8 * N_ALPHA * B if name == "rollout" else 0)
        return {"avg_ms": avg, "launches": cnt, "share_of_step": tot / ms_prof, "algorithmic_bytes": alg,
                "achieved_GBps": alg / avg / 1e6}
    kern = {k: kinfo(k) for k in ("linearize", "backward", "rollout")}
    kern["init_rollout_ms"] = ktimes["init_rollout"][0] / max(1, ktimes["init_rollout"][1])
    kern["other_ms_per_iteration"] = ktimes["other"][0] / max(1, ktimes["other"][1])
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    # command (profiles/traffic.json, written by scripts/ncu_summary.py --traffic); null when absent
    traffic, traffic_all = {}, {}
    try:
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = traffic_all.get(f"B{B}", {})
    except Exception:
        pass
    dom = max(("linearize", "backward", "rollout"), key=lambda k: ktimes[k][0])
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["achieved_GBps"] / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
            "fp64_pipe_active_pct": traffic.get("fp64_pipe_active_pct", {}).get(dom),
            "note": "rollout is FP64-pipe/latency bound at this batch, not HBM bound (DESIGN.md section 4; "
                    "fp64_pipe_active_pct = ncu sm__pipe_fp64_cycles_active of the committed capture, averaged over the "
                    "kernel class incl. the lone-warp alpha=0 rollout; 100 % is out of reach for this instruction mix: "
                    "three-register DFMAs issue at 3 cycles); "
                    "see roofline_backward for the HBM-bound kernel north_star names"}
    roof_b = {"bound": "hbm", "kernel": "backward", "achieved": kern["backward"]["achieved_GBps"], "peak": peak,
              "unit": "GB/s", "frac": kern["backward"]["achieved_GBps"] / peak, "traffic": traffic.get("backward")}

    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ilqr_oracle as O
        threads = O.max_threads()
        p = O.make_problem("ua", "rk4", N_H, DT, UA["Q"], UA["R"], UA["Q_f"], UA["x_target"], UA["phys"], tol=0.0,
                           maxiter=ITERS, n_alpha=N_ALPHA)
        ns = args.cpu_sample
        O.optimize_batch(p, x0_host[:threads], np.zeros((threads, 1, N_H)), nthreads=threads)
        c0 = time.perf_counter()
        r = O.optimize_batch(p, x0_host[:ns], np.zeros((ns, 1, N_H)), nthreads=threads)
        cdt = time.perf_counter() - c0
        cpu = {"value": float(r["iters"].sum()) / cdt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {ns} trajectories of the batch x {ITERS} iterations, {cdt:.1f} s wall on {threads} threads"}

    value = units / (ms * 1e-3)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
           "traj_iterations_per_step": units / args.steps,
           "e2e": None if args.skip_e2e else {"value": units_e / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                              "d2h_bytes_per_step": int(d2h), "ms_per_step": te / args.steps * 1e3},
           "gpu_launches": int(launches), "roofline": roof, "roofline_backward": roof_b, "kernels": kern,
           "large_batch": large, "cpu_baseline": cpu, "clocks": clocks}
    if large:
        large["roofline_backward"]["peak"] = peak
        large["roofline_backward"]["frac"] = large["roofline_backward"]["achieved"] / peak
        large["roofline_backward"]["algorithmic_bytes"] = BYTES["backward"] * N_H * large["batch"]
        large["roofline_backward"]["traffic"] = traffic_all.get(f"B{large['batch']}", {}).get("backward")
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
