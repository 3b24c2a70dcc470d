"""FP64 instruction / flop counts per loop of the hot kernels, from the SASS of a built library (DFMA = 2 flop,
DMUL / DADD = 1).  bench.py's FLOPS table comes from here:

    python scripts/sass_flops.py [path/to/libilqr_b200.so] [kernel-name-substring ...]

For every kernel whose mangled name contains one of the substrings (default: the f64 rk4 UA-double-pendulum hot kernels)
it prints the whole-kernel counts and the counts of every loop (backward branch) with more than 20 DFMAs."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "iterative-linear-quadratic-regulator_b200", "libilqr_b200.so")
    keys = sys.argv[2:] or ["fused_backward_kernelINS_17DoublePendulumSysIdLi1EEENS_8QuadCostIdLi4ELi1EEELi2Ed",
                            "rollout_kernelINS_17DoublePendulumSysIdLi1EEENS_8DiagCostIdLi4ELi1EEELi2Ed",
                            "commit_linearize_kernelINS_17DoublePendulumSysIdLi1EEELi2Ed",
                            "backward_kernelINS_8QuadCostIdLi4ELi1EEEd", "backward_n4m1_lanes_kernelId"]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    f, body = None, collections.defaultdict(list)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            f = m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and f:
            body[f].append((int(m.group(1), 16), m.group(2)))

    def opcode(text):
        t = text.split()
        return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]

    for name, ins in body.items():
        if not any(k in name for k in keys):
            continue
        addr = {a: i for i, (a, _) in enumerate(ins)}
        tot = collections.Counter(opcode(t) for _, t in ins)
        print(name[:110])
        print(f"   whole kernel: {len(ins)} instructions, DFMA {tot['DFMA']} DMUL {tot['DMUL']} DADD {tot['DADD']}")
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                c = collections.Counter(opcode(x) for _, x in ins[addr[tgt]:i + 1])
                if c["DFMA"] > 20:
                    n64 = c["DFMA"] + c["DMUL"] + c["DADD"]
                    print(f"   loop of {i + 1 - addr[tgt]:5d} instructions: DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} "
                          f"(FP64 {n64}) MUFU {c['MUFU']} -> {2 * c['DFMA'] + c['DMUL'] + c['DADD']} flop")


if __name__ == "__main__":
    main()
