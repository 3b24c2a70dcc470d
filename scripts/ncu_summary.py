#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<round>_<what>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    # shared memory: wavefronts (bank-conflict replays included), their share of the LSU data pipe, bank conflicts
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def traffic(rep, key, out):
    """average dram read+write bytes per launch of each kernel class -> out[key] (JSON, merged)"""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    # one bucket per iLQR iteration: K1 (linearize) opens it; the rollout waves (and the two K2 kernels of a large
    # batch) of an iteration are summed, matching bench.py's per-iteration kernel times
    buckets = []
    pipe = {}      # kernel class -> [sum of duration * fp64 pipe %, sum of duration]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        cls = "rollout" if "rollout" in name else "linearize" if "linearize" in name else "backward" if "backward" in name else None
        if cls is None:
            continue
        try:
            dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            pct = float(r[idx["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
            acc_p = pipe.setdefault(cls, [0.0, 0.0])
            acc_p[0] += dur * pct
            acc_p[1] += dur
        except (KeyError, ValueError):
            pass
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[k]].replace(",", "")) * scale[units[idx[k]]]
        # an iteration opens with K1, or -- K1 fused into the scan -- with the fused backward kernel
        if cls == "linearize" or (cls == "backward" and "fused" in name):
            buckets.append({})
        if buckets:
            buckets[-1][cls] = buckets[-1].get(cls, 0.0) + tot
    full = [b for b in buckets if {"backward", "rollout"} <= set(b)]
    classes = [c for c in ("linearize", "backward", "rollout") if full and all(c in b for b in full)]
    acc = {c: [b[c] for b in full] for c in classes}
    d = json.load(open(out)) if os.path.exists(out) else {}
    d[key] = {c: sum(v) / len(v) for c, v in acc.items()}
    # duration-weighted FP64 pipe utilisation per kernel class (a utilisation read under the profiler, not a timing)
    d[key]["fp64_pipe_active_pct"] = {c: v[0] / v[1] for c, v in pipe.items() if v[1] > 0}
    d[key]["source"] = os.path.basename(rep) + " (ncu --set full; per iLQR iteration: sum over the launches of a kernel class, averaged over the captured iterations)"
    json.dump(d, open(out, "w"), indent=1)
    print(d[key])


def main():
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        return traffic(sys.argv[1], sys.argv[3], sys.argv[4])
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none summary of {rep}")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"\n## {name[:150]}")
        for k in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print(f"{k:82s} {r[idx[k]]:>18s} {units[idx[k]]}")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            u = units[idx["dram__bytes_read.sum"]]
            print(f"{'traffic = dram read + write':82s} {rd + wr:18.3f} {u}")
        except Exception:
            pass


if __name__ == "__main__":
    main()
