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
# algorithmic bytes per trajectory-timestep, FP64.  The fused K1+K2 kernel reads
# x_t,u_t from the accepted candidate (40 B), commits them into the nominal (40 B) and writes K_t,k_t (40 B).
BYTES = {"linearize": 240, "backward": 240, "rollout": 120, "fused_backward": 120}
# FP64 flop per trajectory-timestep (rollout: per step size), DFMA = 2, DMUL/DADD = 1, counted from the SASS of the
# f64 / rk4 / UA-double-pendulum kernels by scripts/sass_flops.py: step_jac 485+150+59, riccati_step 246+10+26,
# rollout step 229+61+24 instructions
FLOPS = {"linearize": 1179, "backward": 528, "rollout": 543}
# the same counts for the kernels of batches of TRIG_TABLE_MIN trajectories or more, whose sines and cosines come from the
# shared-memory table (sincos_tab, csrc/ilqr_systems.cuh): step_jac 421+174+75, rollout step 165+85+40 instructions
FLOPS_TABLE = {"linearize": 1091, "backward": 528, "rollout": 455}
TRIG_TABLE_MIN = int(os.environ.get("ILQR_TRIG_TABLE_MIN", "8192"))


def flops_for(B):
    return FLOPS_TABLE if B >= TRIG_TABLE_MIN else FLOPS


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
            "l2_policy": "working set per step (X,U,K,k + 10 candidate slabs = 1.0 GB) exceeds the 126 MB L2; "
                         "no explicit flush"}


def shard_x0(count, rank, world, seed):
    """rank's contiguous shard of a global batch of count*world seeded initial states"""
    return np.ascontiguousarray(cfg2_x0(count * world, seed=seed)[rank * count:(rank + 1) * count])


def evaluated_rollouts(idx, iters, waves, n_first, n_alpha):
    """(step size, trajectory) rollouts the line-search schedule had to evaluate in the traced solve: idx (B, maxiter) =
    accepted try index per iteration (-1 none).  Lazy schedule: every wave up to the one holding the accepted index;
    eager: the first wave, plus the deferred step sizes where none of the first was accepted (speculative extra
    rollouts are not counted)."""
    total = 0
    col = np.arange(idx.shape[1])[None, :]
    ran = col < iters[:, None]
    w = np.where(idx < 0, n_alpha - 1, idx)
    if waves:
        hi = np.cumsum(waves)
        hi[-1] = max(hi[-1], n_alpha)
        upto = hi[np.searchsorted(hi, w + 1)]
    else:
        upto = np.where(w < n_first, n_first, n_alpha)
    total = int(upto[ran].sum())
    return total


def kernel_report(sol, torch, B, steps, peak_hbm, peak_fp64, fused):
    """Per-kernel times (CUDA events chained between the launches of ilqr_solve, on its stream) of `steps` fresh solves,
    and the rooflines: FP64 flop/s against the measured DFMA peak for the pipe-bound kernels, algorithmic bytes/s
    against the measured HBM peak for every kernel."""
    sol.set_profiling(True)
    sol.enable_trace()
    rollouts, units = 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sol.reset_state()
        sol._U.zero_()
        units += sol.solve_device(sync=True)
        idx, _ = sol.trace_arrays()
        rollouts += evaluated_rollouts(idx, sol._iters.cpu().numpy(), sol.linesearch_waves(), sol.first_wave(), N_ALPHA)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kt = sol.kernel_times()
    sol.set_profiling(False)
    nit = max(1, kt["backward"][1] if fused else kt["linearize"][1])         # iLQR iterations profiled
    per_it_rollouts = rollouts / nit
    kern = {}
    FLOPS = flops_for(B)
    for name in ("linearize", "backward", "rollout"):
        tot, cnt = kt[name]
        if fused and name == "linearize":
            continue
        avg = tot / nit
        if name == "rollout":
            flop = FLOPS["rollout"] * N_H * per_it_rollouts
            alg = BYTES["rollout"] * N_H * B + 8 * N_ALPHA * B
        elif name == "backward" and fused:
            flop = (FLOPS["linearize"] + FLOPS["backward"]) * N_H * B
            alg = BYTES["fused_backward"] * N_H * B
        else:
            flop = FLOPS[name] * N_H * B
            alg = BYTES[name] * N_H * B
        kern[name] = {"ms_per_iteration": avg, "launches": cnt, "share_of_step": tot / ms, "algorithmic_bytes": alg,
                      "achieved_GBps": alg / avg / 1e6, "hbm_frac": alg / avg / 1e6 / peak_hbm,
                      "algorithmic_flop": flop, "achieved_TFLOPs": flop / avg / 1e9,
                      "fp64_frac": flop / avg / 1e9 / peak_fp64 if peak_fp64 else None}
    kern["fused_linearize_backward"] = bool(fused)
    kern["rollouts_evaluated_per_iteration"] = per_it_rollouts
    kern["init_rollout_ms"] = kt["init_rollout"][0] / max(1, kt["init_rollout"][1])
    kern["other_ms_per_iteration"] = kt["other"][0] / nit
    return kern, units / (ms * 1e-3)


def make_solver(torch, iLQR, sysm, x0, fused=None):
    """device-resident solver on the bench workload; fused=False builds it on the two-kernel K1/K2 path"""
    old = os.environ.get("ILQR_FUSED")
    if fused is False:
        os.environ["ILQR_FUSED"] = "0"
    try:
        U0 = torch.zeros((1, N_H), dtype=torch.float64, device="cuda")
        return iLQR(sysm, T_H, torch.as_tensor(x0).cuda(), U0, tol=0.0, maxiter=ITERS, verbose=False, n_alpha=N_ALPHA)
    finally:
        if fused is False:
            if old is None:
                os.environ.pop("ILQR_FUSED", None)
            else:
                os.environ["ILQR_FUSED"] = old


def timed_steps(torch, dist, sol, steps, warmup, world, barrier):
    """`steps` fresh solves, device resident, nothing host-synchronous inside the loop: the solve is enqueued with
    sync=False, the iteration counts accumulate on the device, and (N > 1) the per-shard cost/status all-gather -- the
    path's only exchange -- is issued asynchronously so that it runs under the next step's first kernels."""
    B = sol.B
    units_dev = torch.zeros((), dtype=torch.int64, device="cuda")
    bufs = None
    if world > 1:
        bufs = [(torch.empty((2, B), dtype=torch.float64, device="cuda"),
                 torch.empty((world, 2, B), dtype=torch.float64, device="cuda"), [None]) for _ in range(2)]

    def step(i):
        sol.reset_state()
        sol._U.zero_()
        sol.solve_device(sync=False)
        units_dev.add_(sol._iters.sum())
        if world > 1:
            pack, gathered, work = bufs[i & 1]
            if work[0] is not None:
                work[0].wait()                       # the all-gather that last used this pair (two steps ago)
            pack[0].copy_(sol._cost)
            pack[1].copy_(sol._status)
            work[0] = dist.all_gather_into_tensor(gathered, pack, async_op=True)

    for i in range(warmup):
        step(i)
    if world > 1:
        for _, _, work in bufs:
            if work[0] is not None:
                work[0].wait()
    units_dev.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(steps):
        step(i)
    if world > 1:
        for _, _, work in bufs:
            if work[0] is not None:
                work[0].wait()
    e1.record()
    barrier()
    t1 = time.time()
    return e0.elapsed_time(e1), int(units_dev.item()), t0, t1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="trajectories per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="trajectories per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="skip the host-in/host-out arm")
    ap.add_argument("--large-batch", type=int, default=131072,
                    help="also time a few steps at this batch PER GPU (config 5's shard: 8 x 131072 = 1M trajectories), "
                         "where the passes are throughput bound; reported under 'large_batch'.  0 disables.")
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
    # stdout carries exactly ONE line, the JSON: everything else written to fd 1 while the bench runs (NCCL's version
    # banner, library chatter) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # The path's only exchange is a 64 KB all-gather per step, issued asynchronously.  An NCCL kernel that waits for
        # a peer spins on the SMs it occupies; capped to ONE CTA it cannot take more than one SM from the latency-bound
        # solve kernels it runs beside (default channel counts cost 14 % of the step at N=2).
        os.environ.setdefault("NCCL_MAX_CTAS", "1")
        os.environ.setdefault("NCCL_MIN_CTAS", "1")
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = 1
            opts.config.min_ctas = 1
        except Exception:
            opts = None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    from class_files.iLQR_class import iLQR
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum
    from class_files import _cabi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sysm = MyUADoublePendulum(dt=DT, x_target=np.array(UA["x_target"]), Q=np.diag(UA["Q"]), R=np.diag(UA["R"]),
                              Q_f=np.diag(UA["Q_f"]), integrator="rk4", **UA["phys"])
    B = args.batch
    x0_host = shard_x0(B, rank, world, seed=0)
    # ---- device-resident arm: torch tensors in, nothing crosses PCIe in the timed region --------------
    sol = make_solver(torch, iLQR, sysm, x0_host)
    fused = os.environ.get("ILQR_FUSED", "") != "0"
    l0 = None
    sampler = ClockSampler(local) if rank == 0 else None
    # (launch count of the timed steps only: read before and after)
    timed_steps(torch, dist, sol, 0, args.warmup, world, barrier)
    l0 = sol.launches()
    ms, units, t0, t1 = timed_steps(torch, dist, sol, args.steps, 0, world, barrier)
    launches = sol.launches() - l0
    clocks = sampler.stop(t0, t1) if sampler else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "6650 GB/s (fallback)"
    # FP64 FMA peak of THIS device, measured now by the library's DFMA probe (ilqr_fp64_peak)
    tf = __import__("ctypes").c_double(0.0)
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    lib = _cabi.load()
    _cabi.check(lib.ilqr_fp64_peak(__import__("ctypes").byref(tf), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))
    peak_fp64 = float(tf.value)

    # per-kernel breakdown: a few more steps with CUDA events chained between the launches (kept out of the timed
    # region above: the event records cost ~3 % at this batch)
    kern, _ = kernel_report(sol, torch, B, max(1, min(args.steps, 10)), peak, peak_fp64, fused)

    # ---- end-to-end arm: pinned host arrays in, host arrays out through the public API ---------------------
    # optimize_trajectory_async() keeps two result slots: while the 82 MB of X, U of solve i travel to the host on a
    # copy stream, solve i+1 (its own H2D included) already runs.  Every step's H2D and D2H is inside the timed region;
    # `e2e_blocking` is the same loop with optimize_trajectory(), i.e. one solve at a time.
    te = tb = float("nan")
    units_e = units_b = h2d = d2h = 0
    sol_h = None
    if not args.skip_e2e:
        x0_pin = torch.as_tensor(x0_host).pin_memory()
        U_pin = torch.zeros((1, N_H), dtype=torch.float64).pin_memory()
        sol_h = iLQR(sysm, T_H, x0_pin, U_pin, tol=0.0, maxiter=ITERS, verbose=False, n_alpha=N_ALPHA)

        def submit():
            sol_h.x_0 = x0_pin                    # H2D (pinned, asynchronous)
            sol_h.U = U_pin                       # H2D
            sol_h.reset_state()
            return sol_h.optimize_trajectory_async()

        def pipelined(n):
            got, pend = 0, None
            for _ in range(n):
                nxt = submit()
                if pend is not None:
                    X, U, cost = pend.result()    # D2H of X, U, cost of the previous solve has landed
                    got += pend.total_iterations
                pend = nxt
            X, U, cost = pend.result()
            return got + pend.total_iterations, X, U, cost

        pipelined(max(3, args.warmup))
        barrier()
        te0 = time.perf_counter()
        units_e, X, U, cost = pipelined(args.steps)
        barrier()
        te = time.perf_counter() - te0
        h2d = x0_pin.numel() * 8 + U_pin.numel() * 8
        d2h = X.nbytes + U.nbytes + cost.nbytes + 8

        def blocking():
            sol_h.x_0 = x0_pin
            sol_h.U = U_pin
            sol_h.reset_state()
            X, U, cost = sol_h.optimize_trajectory()
            return sol_h.total_iterations

        for _ in range(3):
            blocking()
        barrier()
        tb0 = time.perf_counter()
        for _ in range(max(3, args.steps // 2)):
            units_b += blocking()
        barrier()
        tb = time.perf_counter() - tb0

    # ---- the same solve at config 5's per-GPU shard size (HBM/FP64-throughput-bound regime) -----------------
    large = None
    if args.large_batch > 0:
        del sol_h, sol
        torch.cuda.empty_cache()
        BL = args.large_batch
        sol_l = make_solver(torch, iLQR, sysm, shard_x0(BL, rank, world, seed=1))
        lms, lunits, _, _ = timed_steps(torch, dist, sol_l, 3, 2, world, barrier)
        lkern, _ = kernel_report(sol_l, torch, BL, 2, peak, peak_fp64, fused)
        large = {"batch_per_gpu": BL, "global_batch": BL * world, "steps": 3, "ms_per_step": lms / 3, "units": lunits,
                 "unit": UNIT, "line_search": "lazy waves of 2,2,2,4 step sizes over compacted lists", "kernels": lkern}
        # north_star's evidence for the backward pass is its HBM fraction: with K1 fused into it the pass no longer
        # streams A_t, B_t through HBM at all, so the two-kernel scan is timed once more on the same data (N=1 only)
        if world == 1 and fused:
            del sol_l
            torch.cuda.empty_cache()
            sol_u = make_solver(torch, iLQR, sysm, shard_x0(BL, rank, world, seed=1), fused=False)
            timed_steps(torch, dist, sol_u, 1, 1, world, barrier)
            ukern, _ = kernel_report(sol_u, torch, BL, 2, peak, peak_fp64, False)
            large["two_kernel_path"] = {k: ukern[k] for k in ("linearize", "backward")}
            del sol_u

    if world > 1:
        vals = [ms, te, tb, float(units), float(units_e), float(units_b)]
        if large:
            vals += [large["ms_per_step"], float(large["units"])]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, te, tb = float(tmax[0]), float(tmax[1]), float(tmax[2])
        units, units_e, units_b = int(tsum[3]), int(tsum[4]), int(tsum[5])
        if large:
            large["ms_per_step"], large["units"] = float(tmax[6]), int(tsum[7])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if large:
        large["value"] = large.pop("units") / (large["ms_per_step"] * 3 * 1e-3)
        large["workload"] = (f"cfg5 shard: {large['batch_per_gpu']} trajectories per GPU x {world} GPU(s) = "
                             f"{large['global_batch']} (BASELINE configs[4] is 1M over 8), N={N_H}, {ITERS} iterations per step")

    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    # command (profiles/traffic.json, written by scripts/ncu_summary.py --traffic); null when absent
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tB = traffic.get(f"B{B}", {})
    dom = max((k for k in ("linearize", "backward", "rollout") if k in kern), key=lambda k: kern[k]["share_of_step"])
    kd = kern[dom]
    pipe_bound = dom == "rollout" or fused
    roof = {"kernel": dom, "peak_source": peak_src, "traffic": tB.get(dom),
            "hbm": {"achieved": kd["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kd["hbm_frac"]},
            "fp64": {"achieved": kd["achieved_TFLOPs"], "peak": peak_fp64, "unit": "TFLOP/s", "frac": kd["fp64_frac"],
                     "peak_source": "ilqr_fp64_peak(): DFMA chains on every SM, measured in this run",
                     "flop_per_rollout_step": flops_for(B)["rollout"]},
            "pipe_active_pct_from_profiles": tB.get("fp64_pipe_active_pct", {}).get(dom),
            "note": "the rollout is bound by the FP64 pipe, not by HBM (DESIGN.md section 4): frac = FP64 flop/s (DFMA = 2, "
                    "counted from the kernel's SASS, x rollouts the schedule evaluated) / measured DFMA peak; the HBM figures "
                    "are kept beside it; pipe_active_pct_from_profiles is ncu's number in the committed capture, not "
                    "measured in this run"}
    if pipe_bound:
        roof.update(bound="fp64", achieved=kd["achieved_TFLOPs"], peak=peak_fp64, unit="TFLOP/s", frac=kd["fp64_frac"])
    else:
        roof.update(bound="hbm", achieved=kd["achieved_GBps"], peak=peak, unit="GB/s", frac=kd["hbm_frac"])
    kb = kern["backward"]
    roof_b = {"kernel": "backward" + (" (fused with linearization)" if fused else ""), "bound": "fp64" if fused else "hbm",
              "achieved": kb["achieved_TFLOPs"] if fused else kb["achieved_GBps"],
              "peak": peak_fp64 if fused else peak, "unit": "TFLOP/s" if fused else "GB/s",
              "frac": kb["fp64_frac"] if fused else kb["hbm_frac"], "traffic": tB.get("backward"),
              "hbm": {"achieved": kb["achieved_GBps"], "peak": peak, "frac": kb["hbm_frac"],
                      "algorithmic_bytes_per_trajectory_step": BYTES["fused_backward" if fused else "backward"]}}
    if large:
        lk = large["kernels"]["backward"]
        large["roofline_backward"] = {"bound": "fp64" if fused else "hbm", "unit": "TFLOP/s" if fused else "GB/s",
                                      "achieved": lk["achieved_TFLOPs"] if fused else lk["achieved_GBps"],
                                      "peak": peak_fp64 if fused else peak,
                                      "frac": lk["fp64_frac"] if fused else lk["hbm_frac"],
                                      "traffic": traffic.get(f"B{large['batch_per_gpu']}", {}).get("backward")}
        if "two_kernel_path" in large:
            u = large["two_kernel_path"]["backward"]
            large["roofline_backward"]["two_kernel_scan_hbm"] = {
                "bound": "hbm", "achieved": u["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": u["hbm_frac"],
                "algorithmic_bytes": u["algorithmic_bytes"],
                "note": "backward_kernel alone on K1's materialised A_t, B_t (ILQR_FUSED=0; ring filled by cp.async.bulk onto "
                        "mbarriers): the >= 60 % of HBM roofline north_star asks of the backward pass.  The peak is the driver's "
                        "measured COPY rate (reads = writes); this kernel reads five times what it writes, so a frac "
                        "at or slightly above 1 means 'as fast as a device-to-device copy moves the same bytes'"}

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
    e2e = None
    if not args.skip_e2e:
        e2e = {"value": units_e / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": te / args.steps * 1e3, "pipeline_depth": 2,
               "api": "iLQR.optimize_trajectory_async(): pinned host x_0, U in; host X, U, cost out; the D2H of solve i "
                      "overlaps solve i+1 on a copy stream",
               "blocking": {"value": units_b / tb, "unit": UNIT, "api": "iLQR.optimize_trajectory(), one solve at a time",
                            "ms_per_step": tb / max(3, args.steps // 2) * 1e3}}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
           "traj_iterations_per_step": units / args.steps, "e2e": e2e,
           "gpu_launches": int(launches), "roofline": roof, "roofline_backward": roof_b, "kernels": kern,
           "large_batch": large, "cpu_baseline": cpu, "clocks": clocks}
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
