#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): single-GPU confirmation of the leaner range-major search kernel and the L1-sized row table
# of the tag kernel, then the first run of the key-range-sharded multi-GPU compare (verify + bench at N=1 and N=2).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SMB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_stripe.py tests/test_gpu_experimental.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
bash scripts/gpu_multi.sh "1 2" ${TAG}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"range_major_kernel" -s 2 -c 1 -f -o gpurun_out/rm_${TAG} \
   python bench.py --workload search --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_rm_${TAG}.err; tail -1 gpurun_out/ncu_rm_${TAG}.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> /dev/null
ls gpurun_out | tail -20
