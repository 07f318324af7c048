#!/bin/bash
# Round 2, GPU call (1 GPU): search pass with the two-level row table -- parity at full size, time, ncu; gather.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 900 python -m pytest tests/test_gpu_search_index.py tests/test_gpu_fullsize.py -q -m gpu -x -k "search or gather or index" 2>&1 | tail -3 | tee gpurun_out/tests_${TAG}.log
for W in search gather; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2> gpurun_out/bench_${W}_${TAG}.err; tail -1 gpurun_out/bench_${W}_${TAG}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${TAG}.json')); print(d['metric'][:40], 'ms %.3f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], d.get('roofline',{}).get('frac'), {k: v for k, v in d.items() if k.startswith('parity') or k=='rounds'})"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"range_major_kernel" -s 2 -c 1 -f -o gpurun_out/rm_${TAG} \
   python bench.py --workload search --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
ls gpurun_out | grep ${TAG}
