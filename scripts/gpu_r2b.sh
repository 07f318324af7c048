#!/bin/bash
# Round 2, GPU call 2: the rewritten stripe pipeline (32-bit sort keys, u16 tags, lean count kernel) as default:
# whole GPU suite incl. the gated tests, A/B of the compare variants, search with the fixed ranges launch, then
# the launch list of the default bench and ncu --set full of the count and tag kernels.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SMB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests_${TAG}.log
timeout 900 python tests/tools/ab_variants.py compare search > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err
grep -v "^\[bench\]" gpurun_out/ab_${TAG}.err | tail -30
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
   > /dev/null 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"join_stripe_kernel|stripe_tag_kernel" -c 2 -f -o gpurun_out/stripe_${TAG} \
   python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_stripe_${TAG}.err; tail -2 gpurun_out/ncu_stripe_${TAG}.err
ls -la gpurun_out | tail -8
