#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): the new exchange paths (shards all-gathered in place, 16-bit partial counters, replicated
# gather candidates) verified and timed at N=2; ncu of the pipelined range-major kernel and the launch list of a search step.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2f}
bash scripts/gpu_multi.sh "2" ${TAG} "both gather search"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"range_major_kernel" -s 2 -c 1 -f -o gpurun_out/rm_${TAG} \
   python bench.py --workload search --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_rm_${TAG}.err; tail -1 gpurun_out/ncu_rm_${TAG}.err
timeout 300 python bench.py --workload compare --steps 10 --warmup 3 > gpurun_out/bench_compare_n1_${TAG}.json 2> gpurun_out/bench_compare_n1_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_compare_n1_${TAG}.json')); print('N=1 compare ms %.3f e2e %.1f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d.get('parity_checked_pairs'))"; grep affinity gpurun_out/bench_compare_n1_${TAG}.err
ls gpurun_out | tail -12
