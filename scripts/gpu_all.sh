#!/bin/bash
# everything on one GPU: test-suite, default bench, extra workloads
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-all}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
timeout 1000 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.2f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.2f"%x["roofline"]["kernel_ms"], "frac %.3f"%x["roofline"]["frac"], "launches", x["gpu_launches"], "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0))
show(d); show(d["sketch"])
PY
for W in search gather; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> gpurun_out/bench_${W}_${TAG}.err
  tail -2 gpurun_out/bench_${W}_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'], d['ms_per_step'], 'ms', {k:v for k,v in d.items() if k in ('rounds','subjects_per_s','algorithmic_GBps','gpu_launches')})"
done
