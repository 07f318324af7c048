#!/bin/bash
# Round 2, last GPU call (1 GPU): count kernel with 32-bit shared addresses for its increments and one widened tag address per
# element -- whole GPU suite, smoke, A/B of the switches in one process (identical matrices asserted), the default bench
# line, search / gather lines of the final tree, launch list + ncu of the count kernel.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 900 python -m pytest tests -q -m gpu -x --durations=3 2>&1 | tail -8 | tee gpurun_out/tests_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tests/tools/ab_variants.py compare > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err; grep "^compare" gpurun_out/ab_${TAG}.err | cut -c1-220
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -2 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.3f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.3f"%x["roofline"]["kernel_ms"], "launches", x["gpu_launches"], "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0), x.get("clocks"),
          {k: v for k, v in x.items() if k.startswith("parity")})
show(d); show(d["sketch"])
PY
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
   --log-file gpurun_out/launches_compare_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel" -s 3 -c 1 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
for W in search gather; do
  timeout 400 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> gpurun_out/bench_${W}_${TAG}.err; tail -1 gpurun_out/bench_${W}_${TAG}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'][:40], 'ms %.3f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], d.get('roofline',{}).get('frac'), {k: v for k, v in d.items() if k.startswith('parity') or k=='rounds'})"
done
ls gpurun_out | grep ${TAG}
