#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed).

    python scripts/summarize_ncu.py launches <launches.csv> <out.md> ["title"]
        per-kernel totals / shares / DRAM bytes of an `ncu --metrics gpu__time_duration.sum,dram__bytes_*` launch list
    python scripts/summarize_ncu.py step <launches.csv> <first kernel of a step> <k-th occurrence> <out.md>
        the launches of ONE step in order (from the k-th launch of the named kernel to the next one)
    python scripts/summarize_ncu.py full <report.ncu-rep> <out.txt>
        the metrics of an `ncu --set full` capture that the design discussion uses, per captured kernel
"""
import collections
import csv
import re
import subprocess
import sys


def read_launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, mi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Grid Size")
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name).replace("void ", "")
        d = per.setdefault(r[0], {"name": name, "grid": r[gi]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    return list(per.values())


def launches(path, out, title):
    per = read_launches(path)
    agg = collections.OrderedDict()
    for d in per:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0) / 1e6
        a[2] += (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write("# %s\n\nncu launch list (`gpu__time_duration.sum`, `--clock-control none`); per-launch times are cold-cache and serialised: "
                 "compare shares, not absolutes.\n\n| kernel | launches | total ms | avg ms | share | DRAM MB / launch |\n|---|---:|---:|---:|---:|---:|\n" % title)
        for k, a in agg.items():
            fh.write("| `%s` | %d | %.3f | %.4f | %.1f%% | %.1f |\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2] / a[0]))


def step(path, first, kth, out):
    per = read_launches(path)
    idx = [i for i, d in enumerate(per) if d["name"].endswith(first)]
    s, e = idx[int(kth)], (idx[int(kth) + 1] if int(kth) + 1 < len(idx) else len(per))
    with open(out, "w") as fh:
        fh.write("| kernel | grid | ms | DRAM MB |\n|---|---|---:|---:|\n")
        tot = 0.0
        for d in per[s:e]:
            ms = d.get("gpu__time_duration.sum", 0) / 1e6
            tot += ms
            fh.write("| `%s` | %s | %.4f | %.1f |\n" % (d["name"], d["grid"], ms, (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6))
        fh.write("| **sum** | | **%.3f** | |\n" % tot)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_requests_srcunit_tex_op_red.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write("ncu --set full --clock-control none  (%s)\n" % rep.split("/")[-1])
        for vals in rows[2:]:
            fh.write("\n== %s\n" % vals[hdr.index("Kernel Name")][:110])
            for k in WANT:
                if k in hdr:
                    fh.write("%-82s %-14s %s\n" % (k, units[hdr.index(k)], vals[hdr.index(k)]))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
    elif mode == "step":
        step(*sys.argv[2:6])
    else:
        full(sys.argv[2], sys.argv[3])
