#!/usr/bin/env python3
"""Summarise an .ncu-rep (captured with --set full --import-source on) into the two text files kept
under profiles/: key raw metrics per kernel, and the hottest source lines by executed instructions
and by stall samples.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_r01_<name>"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__grid_size launch__block_size
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio sm__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum smsp__warps_active.avg.per_cycle_active
smsp__warps_eligible.avg.per_cycle_active sm__cycles_elapsed.avg sm__cycles_active.avg""".split()


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, outp = sys.argv[1], sys.argv[2]
    raw = ncu(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(outp + "_summary.txt", "w") as f:
        for r in data:
            d = dict(zip(hdr, r))
            f.write("kernel: %s grid %s block %s\n" % (d.get("Kernel Name", "?")[:100], d.get("Grid Size"), d.get("Block Size")))
            for k in hdr:
                if k in KEYS or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                    f.write("  %-86s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    # cuda,sass view: per file a table whose rows are either a source line (non-empty "Line No", metrics
    # aggregated over its SASS) or one SASS instruction (empty "Line No")
    inst = defaultdict(float)
    smp = defaultdict(float)
    head, fname = None, "?"
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            head = r
            continue
        if head is None or len(r) != len(head) or r[0] == "":
            continue
        try:
            ie = float(r[head.index("Instructions Executed")] or 0)
            ws = float(r[head.index("# Samples")] or 0)
        except ValueError:
            continue
        key = "%s:%s  %s" % (fname, r[0], r[1].strip()[:100])
        inst[key] += ie
        smp[key] += ws
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    with open(outp + "_hotlines.txt", "w") as f:
        f.write("total inst %d samples %d\n--- by instructions\n" % (ti, ts))
        for k, v in sorted(inst.items(), key=lambda kv: -kv[1])[:45]:
            f.write("%5.1f%% inst %5.1f%% smp  %s\n" % (100 * v / ti, 100 * smp[k] / ts, k))
        f.write("--- by stall samples\n")
        for k, v in sorted(smp.items(), key=lambda kv: -kv[1])[:45]:
            f.write("%5.1f%% inst %5.1f%% smp  %s\n" % (100 * inst[k] / ti, 100 * v / ts, k))


if __name__ == "__main__":
    main()
