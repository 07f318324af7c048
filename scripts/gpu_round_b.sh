#!/bin/bash
# GPU call B: new tests (join, counter-gather port, hypothesis protein, ANI), A/B of the compare
# algorithms on the 10k workload, launch list of the default bench
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1p}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_counter_gather_port.py tests/test_gpu_hypothesis.py \
    tests/test_gpu_gather_report.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -q -m gpu --durations=6 2>&1 | tail -30 | tee gpurun_out/tests_${TAG}.log
for ALGO in join tile; do
  SMB_COMPARE_ALGO=$ALGO timeout 300 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_compare_${ALGO}_${TAG}.json 2> gpurun_out/bench_compare_${ALGO}_${TAG}.err
  tail -2 gpurun_out/bench_compare_${ALGO}_${TAG}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_compare_${ALGO}_${TAG}.json"))
print("${ALGO}", "value %.4g"%d["value"], "ms %.2f"%d["ms_per_step"], "e2e %.4g (%.1f ms)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), "kernel_ms %.2f"%d["roofline"]["kernel_ms"], d["roofline"]["algorithm"], "launches", d["gpu_launches"])
PY
done
timeout 300 python bench.py --workload compare --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_compare_auto_${TAG}.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_compare_auto_${TAG}.json')); print('auto', d['ms_per_step'], d['roofline']['algorithm'])"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 1 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/launches_${TAG}.err
tail -2 gpurun_out/launches_${TAG}.err; ls -la gpurun_out | tail -8
