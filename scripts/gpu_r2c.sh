#!/bin/bash
# Round 2, GPU call 3: leaner stripe count kernel, range-major search pass, new configs[3]/[4] workloads, bench with
# in-line parity checks.  Whole GPU suite, the three bench workloads, launch lists, ncu --set full of the two hot kernels.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2c}
nproc; free -g | sed -n 2p
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SMB_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -q -m gpu -x --durations=6 2>&1 | tail -14 | tee gpurun_out/tests_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
def show(x):
    print(x["metric"], "value %.4g"%x["value"], "ms %.3f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]),
          "kernel_ms %.3f"%x["roofline"]["kernel_ms"], "frac %.3f"%x["roofline"]["frac"], "launches", x["gpu_launches"], "cpu %.4g"%x.get("cpu_baseline",{}).get("value",0), x.get("clocks"),
          {k: v for k, v in x.items() if k.startswith("parity")})
show(d); show(d["sketch"]); print(json.dumps(d["sketch"].get("variants")))
PY
for W in search gather; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> gpurun_out/bench_${W}_${TAG}.err; tail -2 gpurun_out/bench_${W}_${TAG}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'], 'ms %.3f'%d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline'), {k: v for k, v in d.items() if k.startswith('parity') or k=='rounds'})"
done
timeout 600 python tests/tools/ab_variants.py compare search gather > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err
grep -v "^\[bench\]" gpurun_out/ab_${TAG}.err | tail -30
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches_search_${TAG}.csv python bench.py --workload search --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel" -c 1 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_stripe_${TAG}.err; tail -1 gpurun_out/ncu_stripe_${TAG}.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"range_major_kernel" -s 2 -c 1 -f -o gpurun_out/rm_${TAG} \
   python bench.py --workload search --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_rm_${TAG}.err; tail -1 gpurun_out/ncu_rm_${TAG}.err
ls -la gpurun_out | tail -12
