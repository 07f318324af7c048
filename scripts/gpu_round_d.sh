#!/bin/bash
# GPU call D: join kernel variants (0 = global walk, 1 = shared-memory tile) on the 10k matrix
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1s}
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "join" 2>&1 | tail -3
for V in 0 1; do
  SMB_JOIN_VARIANT=$V timeout 200 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_join_v${V}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_join_v${V}_${TAG}.json')); print('variant ${V}: ms %.2f kernel_ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:join_count -s 3 -c 1 \
    -o gpurun_out/prof_join_${TAG} -f python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof_join_${TAG}.err
tail -1 gpurun_out/prof_join_${TAG}.err
