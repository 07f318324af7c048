#!/bin/bash
# one GPU, everything: full GPU suite, default bench (both headline workloads), widened-row measurements,
# search / gather workloads, launch list of the default bench
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1t}
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.2f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.2f"%x["roofline"]["kernel_ms"], "frac %.3f"%x["roofline"]["frac"], "launches", x["gpu_launches"], "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0), x.get("clocks"))
show(d); show(d["sketch"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cat gpurun_out/bench_ref_${TAG}.json | cut -c1-400
timeout 400 python tests/tools/bench_extra.py > gpurun_out/extra_${TAG}.json 2> gpurun_out/extra_${TAG}.err; tail -2 gpurun_out/extra_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/extra_${TAG}.json')); print(json.dumps(d['files'])); print(json.dumps(d['sigs']))"
for W in search gather; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'], d['ms_per_step'], 'ms')"
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/launches_${TAG}.err
tail -1 gpurun_out/launches_${TAG}.err
