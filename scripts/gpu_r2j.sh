#!/bin/bash
# Round 2, GPU call (1 GPU): the new shard / take_rows / device-query GPU tests, gather after the faster intersect kernel.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 900 python -m pytest tests/test_gpu_stripe.py tests/test_gpu_api.py tests/test_gpu_counter_gather_port.py tests/test_gpu_gather_report.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
timeout 600 python bench.py --workload gather --steps 10 --warmup 3 > gpurun_out/bench_gather_${TAG}.json 2> gpurun_out/bench_gather_${TAG}.err; tail -1 gpurun_out/bench_gather_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_gather_${TAG}.json')); print(d['metric'][:40], 'ms %.3f'%d['ms_per_step'], {k: v for k, v in d.items() if k.startswith('parity') or k=='rounds'})"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_gather_${TAG}.csv \
   python bench.py --workload gather --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
