#!/bin/bash
# Next round, first (short) GPU call: do the experimental paths give the right answers at all, and how do
# the most promising ones compare with the defaults?  ~6 minutes.  scripts/gpu_next_variants.sh is the
# full validation + A/B + profiles.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
SMB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/experimental_${TAG}.log
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); d=d.get('sketch', d) if sys.argv[2]=='sketch' else d
print(sys.argv[3], 'ms %.2f'%d['ms_per_step'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'] if 'roofline' in d else '', 'e2e %.1f ms'%d['e2e']['ms_per_step'] if 'e2e' in d else '', d.get('index',''))" "$@"; }
for L in plain stripe_upper; do
  SMB_JOIN_LAYOUT=$L SMB_JOIN_SORT=$([ $L = plain ] && echo full || echo low32) timeout 200 python bench.py --workload compare --steps 5 --warmup 3 \
      --no-cpu-baseline > gpurun_out/q_compare_${L}_${TAG}.json 2> /dev/null && show gpurun_out/q_compare_${L}_${TAG}.json compare "compare $L:"
done
for F in 0 1; do
  SMB_SKETCH_FUSED=$F timeout 200 python bench.py --workload sketch --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/q_sketch_fused${F}_${TAG}.json 2> /dev/null && show gpurun_out/q_sketch_fused${F}_${TAG}.json sketch "sketch fused=$F:"
done
timeout 400 python bench.py --workload gather --steps 3 --warmup 3 > gpurun_out/q_gather_plain_${TAG}.json 2> /dev/null && show gpurun_out/q_gather_plain_${TAG}.json x "gather plain:"
timeout 400 python bench.py --workload gather --index --steps 3 --warmup 3 > gpurun_out/q_gather_index_${TAG}.json 2> /dev/null && show gpurun_out/q_gather_index_${TAG}.json x "gather index:"
