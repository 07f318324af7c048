#!/bin/bash
# Next round, first GPU call: validate and measure the experimental cluster layout of the inverted
# join (SMB_JOIN_LAYOUT=cluster; default off, logic covered on the CPU by
# tests/test_host_emulation.py::test_join_cluster_layout_matches_oracle).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2a}
# 1. correctness: the join tests with the layout switched on (they compare with the oracle)
SMB_JOIN_LAYOUT=cluster timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu -k "join or compare" 2>&1 | tail -4
# 2. A/B on the 10 000-sketch matrix
for L in plain cluster; do
  SMB_JOIN_LAYOUT=$L timeout 200 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_join_${L}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_join_${L}_${TAG}.json')); print('${L}: ms %.2f kernel_ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
# 2b. row-block passes for the end-to-end path (SMB_COMPARE_PASSES): correctness through the host API, then e2e
SMB_COMPARE_PASSES=8 timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -q -m gpu -k "compare or join" 2>&1 | tail -3
for P in 0 4 8 16; do
  SMB_COMPARE_PASSES=$P timeout 200 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_passes_${P}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_passes_${P}_${TAG}.json')); print('passes ${P}: e2e %.1f ms'%d['e2e']['ms_per_step'])"
done
# 3. where the time goes
SMB_JOIN_LAYOUT=cluster ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum \
    --clock-control none -c 200 --csv --log-file gpurun_out/launches_cluster_${TAG}.csv \
    python bench.py --workload compare --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/launches_cluster_${TAG}.err
tail -1 gpurun_out/launches_cluster_${TAG}.err
