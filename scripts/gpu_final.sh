#!/bin/bash
# last GPU call of the round: smoke, full GPU suite, default bench
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1u}
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/tests_${TAG}.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.2f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.2f"%x["roofline"]["kernel_ms"], "launches", x["gpu_launches"], x.get("clocks"))
show(d); show(d["sketch"])
PY
