#!/bin/bash
# Round 2, GPU call 1: do the switched paths give the right answers on hardware, and which win?
# smoke, the gated parity tests, one-process A/B of every switch (tests/tools/ab_variants.py), then
# ncu --set full of the stripe count kernel and the ranges search kernel.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
SMB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/experimental_${TAG}.log
timeout 900 python tests/tools/ab_variants.py compare sketch search gather > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err
grep -v "^\[bench\]" gpurun_out/ab_${TAG}.err | tail -40
SMB_JOIN_LAYOUT=stripe_upper SMB_JOIN_SORT=low32 timeout 300 ncu --set full --clock-control none --import-source on \
   -k regex:join_stripe_kernel -c 1 -f -o gpurun_out/stripe_${TAG} python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
   > /dev/null 2> gpurun_out/ncu_stripe_${TAG}.err; tail -2 gpurun_out/ncu_stripe_${TAG}.err
SMB_JOIN_LAYOUT=stripe_upper SMB_JOIN_SORT=low32 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches_stripe_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
   > /dev/null 2> /dev/null
ls -la gpurun_out | tail -20
