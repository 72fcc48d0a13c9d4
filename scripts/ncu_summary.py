#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics per launch, stall-reason totals and the
instructions that collect the most stall samples.  Usage: ncu_summary.py REPORT [--out FILE]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[sys.argv.index("--out") + 1], "w") if "--out" in sys.argv else sys.stdout
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {rep}", file=out)
    for r in data:
        print(f"\n## launch: {r[hdr.index('Kernel Name')]}", file=out)
        for k in KEYS:
            if k in hdr:
                print(f"{k:75s} {r[hdr.index(k)]:>20s} {units[hdr.index(k)]}", file=out)
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hi = [i for i, r in enumerate(src) if "Source" in r and "# Samples" in r]
    if hi:
        h = src[hi[0]]
        end = hi[1] - 1 if len(hi) > 1 else len(src)
        body = [r for r in src[hi[0] + 1:end] if len(r) >= len(h)]
        cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = collections.Counter()
        for r in body:
            for i in cols:
                try:
                    tot[h[i]] += int(r[i])
                except ValueError:
                    pass
        s = sum(tot.values()) or 1
        print("\n## warp stall samples, first launch (all instructions)", file=out)
        for k, v in tot.most_common(10):
            print(f"{k:28s} {v:10d} {100 * v / s:5.1f}%", file=out)
        si, ss, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        top = sorted(body, key=lambda r: -int(r[ss]) if r[ss].isdigit() else 0)[:14]
        print("\n## instructions with the most samples (samples, executed, SASS)", file=out)
        for r in top:
            reasons = sorted(((int(r[i]), h[i]) for i in cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
            print(f"{r[ss]:>8s} {r[ei]:>12s}  {r[si].strip()[:64]:64s} {reasons}", file=out)
        ops = collections.Counter()
        for r in body:
            if r[ei].isdigit():
                t = r[si].split()
                if t:
                    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
                    ops[op] += int(r[ei])
        print("\n## executed warp-instructions by opcode", file=out)
        print(", ".join(f"{k}:{v}" for k, v in ops.most_common(14)), file=out)


if __name__ == "__main__":
    main()
