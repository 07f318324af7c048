#!/bin/bash
# Round 2, GPU call 7 (1 GPU): count kernel driven by rem[] (members behind each stream position): stripe tests, bench, launch
# list of a compare step, ncu --set full of the count / tag / rem kernels.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_stripe.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
timeout 600 python bench.py --workload compare --steps 10 --warmup 3 > gpurun_out/bench_compare_n1_${TAG}.json 2> gpurun_out/bench_compare_n1_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_compare_n1_${TAG}.json')); print('N=1 compare ms %.3f kernel %.3f e2e %.1f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']), d.get('parity_checked_pairs'), d['clocks'])"; grep affinity gpurun_out/bench_compare_n1_${TAG}.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel|stripe_tag_kernel|stripe_rem_kernel" -c 3 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_stripe_${TAG}.err; tail -1 gpurun_out/ncu_stripe_${TAG}.err
ls gpurun_out | tail -6
