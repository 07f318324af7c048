#!/bin/bash
# GPU tests + default bench; outputs under gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-check}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
timeout 1000 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.2f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.2f"%x["roofline"]["kernel_ms"], "frac %.3f"%x["roofline"]["frac"], "launches", x["gpu_launches"], "clocks", x["clocks"],
          "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0))
show(d); show(d["sketch"])
PY
