#!/usr/bin/env python
"""Turn the ncu outputs that scripts/gpu_profile.sh left in gpurun_out/ into the tracked
summaries under profiles/ (run here, in the CPU container):

    python scripts/summarize_profile.py r1

writes profiles/<tag>_launches.csv (raw ncu launch list), profiles/<tag>_launches.md (per-kernel
totals and shares), profiles/<tag>_tile.txt / <tag>_hash.txt (selected --set full metrics) and
updates profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def unit_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)


# ---- launch list
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copyfile(src, os.path.join(P, f"{tag}_launches.csv"))
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    agg = collections.OrderedDict()
    dram = collections.OrderedDict()       # per kernel: DRAM bytes (when the capture carried them)
    BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].split("<")[0]
        if r[mi] == "gpu__time_duration.sum":
            agg.setdefault(name, []).append(unit_ms(r[vi], r[ui]))
        elif r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram[name] = dram.get(name, 0.0) + float(r[vi].replace(",", "")) * BYTES.get(r[ui], 1)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none)\n\n")
        fh.write("Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu; per-launch times are "
                 "cold-cache and serialised, compare shares.\n\n| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            fh.write(f"| `{k}` | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {100 * sum(v) / tot:.1f}% |\n")
        if dram:
            fh.write("\nDRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, same capture):\n\n"
                     "| kernel | MB / launch |\n|---|---:|\n")
            for k, b in dram.items():
                fh.write(f"| `{k}` | {b / len(agg[k]) / 1e6:.1f} |\n")
    print(open(os.path.join(P, f"{tag}_launches.md")).read())

# ---- full captures
traffic_path = os.path.join(P, "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for short, key in (("tile", "pairwise_tile_split_kernel"), ("hash", "hash_kmers_kernel")):
    rep = os.path.join(G, f"prof_{short}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_metrics.py"), rep],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag}_{short}.txt"), "w") as fh:
        fh.write(f"# selected metrics of {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n")
        fh.write(out)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    tot_bytes, n = 0.0, 0
    for r in rows[2:]:
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            b += float(r[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
        tot_bytes += b
        n += 1
    # hash: three launches (k=21,31,51) form one pass -> report the sum; tile: one launch
    traffic[key] = tot_bytes if short == "hash" else tot_bytes / max(n, 1)
    traffic[key + "_source"] = f"profiles/{tag}_{short}.txt"
    print(key, "DRAM bytes:", traffic[key])
json.dump(traffic, open(traffic_path, "w"), indent=1)
