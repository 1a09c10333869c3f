#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the committed summaries
under profiles/:  <tag>_launches.csv (+ _launches_summary.txt) and <tag>_<kernel>_full.txt
(raw metrics, instruction mix, hottest source lines).

  python scripts/summarise_profile.py r01 gpurun_out/launches.csv gpurun_out/prof_cons_jac.ncu-rep cons_jac
"""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_no_instruction.pct", "sm__inst_executed.avg.per_cycle_active"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(tag, path):
    out = os.path.join(ROOT, "profiles", tag + "_launches.csv")
    shutil.copy(path, out)
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[ki], r[gi]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write("# grouped by (kernel, grid): the timed step of bench.py is ONE k_cons_jac launch over the whole batch (grid x = 2048;\n")
        f.write("# the quadrotor has no events or linkages, so no k_endpoint); the 8x smaller grids and k_return_head belong to the\n")
        f.write("# chunked host-pointer (e2e) call\n")
        f.write("%8s %12s %12s %7s  %-18s kernel\n" % ("launches", "avg_ns", "total_ns", "share", "grid"))
        for (k, g), v in d.items():
            f.write("%8d %12.0f %12.0f %6.1f%%  %-18s %s\n" % (len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot, g, k[:150]))
    print(open(os.path.join(ROOT, "profiles", tag + "_launches_summary.txt")).read())


def full(tag, rep, name):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, u = raw[0], raw[1]
    lines = ["# ncu --set full --clock-control none --import-source on : %s" % os.path.basename(rep)]
    for r in raw[2:]:
        lines.append("## launch id %s  %s" % (r[h.index("ID")], r[h.index("Kernel Name")][:140]))
        for k in KEYS:
            if k in h:
                lines.append("%-75s %18s %s" % (k, r[h.index(k)], u[h.index(k)]))
    # instruction mix + stall reasons (SASS view)
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-count", "1"]))))
    hh = src[1]
    ix = {k: i for i, k in enumerate(hh)}
    seen, data = set(), []
    for r in src[2:]:  # the CSV repeats the listing once per view: keep each address once
        if len(r) == len(hh) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    ops, stalls, tot = collections.Counter(), collections.Counter(), 0
    for r in data:
        t = r[ix["Source"]].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        tot += n
        for k in hh:
            if k.startswith("stall_") and "Not Issued" not in k:
                stalls[k] += int(r[ix[k]])
    lines.append("## instruction mix (first launch): %d static SASS instructions, %d warp-level instructions executed" % (len(data), tot))
    for op, n in ops.most_common(14):
        lines.append("%-10s %12d %5.1f%%" % (op, n, 100.0 * n / max(tot, 1)))
    st = sum(stalls.values())
    lines.append("## warp stall samples: " + ", ".join("%s %.0f%%" % (k, 100.0 * v / max(st, 1)) for k, v in stalls.most_common(6)))
    # hottest CUDA source lines
    cs = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"]))))
    fil, out = None, []
    for r in cs:
        if len(r) >= 2 and r[0] == "File Path":
            fil = r[1].split("/")[-1]
        elif len(r) > 8 and r[0].isdigit():
            try:
                out.append((int(r[7]), fil, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    t2 = sum(o[0] for o in out)
    lines.append("## hottest source lines (warp-level instructions executed)")
    for n, f, ln, s in sorted(out, reverse=True)[:16]:
        lines.append("%10d %5.1f%%  %s:%d  %s" % (n, 100.0 * n / max(t2, 1), f, ln, s))
    path = os.path.join(ROOT, "profiles", "%s_%s_full.txt" % (tag, name))
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if sys.argv[2] != "-":
        launches(tag, sys.argv[2])
    if len(sys.argv) > 4:
        full(tag, sys.argv[3], sys.argv[4])
