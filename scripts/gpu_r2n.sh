#!/bin/bash
# Round 2, last GPU call (1 GPU): the stream build without the separate descent pass (noticed by the tag kernel, redone by a
# warp per run), 64 x 64 mirror tiles, two count-kernel CTAs per SM -- whole GPU suite, smoke, A/B of the switches in one
# process (identical matrices asserted), the default bench line, launch list + ncu of the changed kernels.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 900 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -10 | tee gpurun_out/tests_${TAG}.log
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
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_compare_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel|stripe_tag_kernel|stripe_fix_kernel|stripe_mirror_kernel" -s 12 -c 4 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
ls gpurun_out | grep ${TAG}
