#!/usr/bin/env python3
"""Turn one `ncu --set full --import-source on` capture of fp_chain2_kernel into the two small text files kept under profiles/:

    <out>_raw.csv           the launch's metrics that the docs quote (duration, instructions, issue rate, stall reasons, DRAM bytes ...)
    <out>_by_function.txt   warp-instructions per unit and average active threads per device function (source-level attribution),
                            the hottest source lines, and where the barrier / long-scoreboard samples sit

usage:  scripts/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/r02_x <units in the profiled launch>
The capture must come from the sources in the working tree (line attribution is by file:line).
"""
import bisect
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum")


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    raw = ncu_csv(rep, "raw")
    hdr, unit_row, val = raw[0], raw[1], raw[-1]
    with open(out + "_raw.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for i, k in enumerate(hdr):
            if k in KEEP or ("issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k) or k in ("Kernel Name", "Block Size", "Grid Size"):
                w.writerow([k, unit_row[i] if i < len(unit_row) else "", val[i]])
    rows = ncu_csv(rep, "source", ("--print-source", "cuda,sass"))
    per, cur, hdr = {}, "", None
    for r in rows:
        if r and r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if r and r[0] == "Line No":
            hdr = r
            ie, it, ism = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            ib, il = hdr.index("stall_barrier"), hdr.index("stall_long_sb")
            continue
        if hdr and r and r[0].isdigit() and len(r) > max(it, ib, il):
            try:
                ln = int(r[0]); v = [int(r[k] or 0) for k in (ie, it, ism, ib, il)]
            except ValueError:
                continue
            a = per.setdefault((cur, ln), [0, 0, 0, 0, 0, r[1]])
            for k in range(5):
                a[k] += v[k]
    funcs = {}
    for f in ("fp_chain2.cuh", "fp_device.cuh"):
        fl = []
        for i, l in enumerate(open(os.path.join(ROOT, "fastp_b200", "csrc", f)).read().split("\n"), 1):
            if l.startswith(("__device__", "__global__", "template")):
                nm = re.findall(r"(\w+)\s*\(", l)
                if nm:
                    fl.append((i, nm[0] if nm[0] != "__launch_bounds__" else "fp_chain2_kernel (body)"))
        funcs[f] = fl
    agg = {}
    for (f, ln), v in per.items():
        if f in funcs:
            st = [x[0] for x in funcs[f]]
            k = bisect.bisect_right(st, ln) - 1
            name = funcs[f][k][1] if k >= 0 else "?"
        else:
            name = f
        a = agg.setdefault(name, [0, 0])
        a[0] += v[0]; a[1] += v[1]
    tot_i = sum(v[0] for v in per.values()); tot_s = sum(v[2] for v in per.values())
    with open(out + "_by_function.txt", "w") as f:
        f.write(f"# {os.path.basename(rep)}: {units:.0f} units in the launch, {tot_i / units:.1f} warp-instructions per unit (source-attributed), {tot_s} stall samples\n")
        f.write("# function                          warp-inst/unit   avg active threads\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            f.write(f"{k:34s} {v[0] / units:10.1f} {v[1] / max(v[0], 1):12.1f}\n")
        f.write("\n# hottest source lines (share of all stall samples | warp-inst/unit | avg threads)\n")
        for (fn, ln), v in sorted(per.items(), key=lambda kv: -kv[1][2])[:25]:
            f.write(f"{fn}:{ln:<5d} {100 * v[2] / max(tot_s, 1):5.2f}%  {v[0] / units:7.2f}  {v[1] / max(v[0], 1):5.1f}  {v[5].strip()[:100]}\n")
        f.write("\n# barrier-stall samples by line (share of all samples)\n")
        for (fn, ln), v in sorted(per.items(), key=lambda kv: -kv[1][3])[:8]:
            f.write(f"{fn}:{ln:<5d} {100 * v[3] / max(tot_s, 1):5.2f}%  {v[5].strip()[:100]}\n")
        f.write("\n# long-scoreboard samples by line\n")
        for (fn, ln), v in sorted(per.items(), key=lambda kv: -kv[1][4])[:8]:
            f.write(f"{fn}:{ln:<5d} {100 * v[4] / max(tot_s, 1):5.2f}%  {v[5].strip()[:100]}\n")


if __name__ == "__main__":
    main()
