#!/bin/bash
# Round 2, GPU call 8 (1 GPU): confirmation of the committed tree the way the driver runs it -- whole GPU suite, smoke, the
# default bench line (--steps 20 --warmup 5), the reference arm, search / gather workloads -- plus the launch lists and ncu
# captures the roofline records are taken from.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -10 | tee gpurun_out/tests_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -2 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.3f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.3f"%x["roofline"]["kernel_ms"], "launches", x["gpu_launches"], "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0), x.get("clocks"),
          {k: v for k, v in x.items() if k.startswith("parity")})
show(d); show(d["sketch"])
PY
for W in search gather; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> gpurun_out/bench_${W}_${TAG}.err; tail -1 gpurun_out/bench_${W}_${TAG}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'][:40], 'ms %.3f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], d.get('roofline',{}).get('frac'), {k: v for k, v in d.items() if k.startswith('parity') or k=='rounds'})"
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2> /dev/null; cut -c1-300 gpurun_out/bench_reference_${TAG}.json
for W in compare sketch search gather; do
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
     --log-file gpurun_out/launches_${W}_${TAG}.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel|stripe_tag_kernel" -c 2 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hash_kmers_fused_kernel" -c 1 -f -o gpurun_out/hash_${TAG} \
   python bench.py --workload sketch --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
ls gpurun_out | tail -16
