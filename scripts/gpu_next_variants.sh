#!/bin/bash
# Next round, first GPU call: validate and measure the experimental layouts of the inverted join that
# are in the tree behind switches (default off; logic covered on the CPU by tests/test_host_emulation.py):
#   SMB_JOIN_LAYOUT=stripe   CTAs own complete result rows in shared memory, no global atomics, fused
#                            finalize, row blocks downloadable as they finish (csrc/join_stripe.cuh);
#                            stripe_upper: only (i, j > i) counted, the rest mirrored tile by tile;
#                            SMB_JOIN_SORT=low32: the stream sorted on the low key words (4 passes) + repair
#   SMB_JOIN_LAYOUT=cluster  related rows at adjacent ranks, warp per element (csrc/join_walk.cuh)
#   SMB_COMPARE_PASSES=k     row-block count passes for the end-to-end path
#   SMB_SKETCH_FUSED=1       sketch: k = 21, 31, 51 in one pass over the bases (csrc/kmer_roll.cuh)
#   SketchSet.build_index()  inverted index of a resident set: search / gather counts for work proportional
#                            to the query (csrc/db_index.cuh); bench.py --workload search|gather --index
#   SMB_SEARCH_LAYOUT=ranges one CTA per key range, query bitmap of the range in shared memory, for the
#                            one-vs-many pass of search / prefetch / gather (csrc/range_search.cuh)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2a}
# 0. the explicit parity tests of the experimental paths (skipped in the default suite)
SMB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -q -m gpu 2>&1 | tail -8
# 1. correctness: the compare / join tests with each layout switched on and the join forced (they
#    compare with the oracle bit for bit); a memcheck run of the stripe kernels on a small matrix
for L in stripe cluster; do
  echo "== tests with SMB_JOIN_LAYOUT=$L"
  SMB_COMPARE_ALGO=join SMB_JOIN_LAYOUT=$L timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py \
      tests/test_gpu_api.py -q -m gpu -k "join or compare" 2>&1 | tail -4
done
SMB_COMPARE_ALGO=join SMB_JOIN_LAYOUT=stripe timeout 600 compute-sanitizer --tool memcheck python -c "
import numpy as np, oracle as orc
from sourmash_b200 import batch as B
from sourmash_b200.synth import synth_sketches
h, off = synth_sketches(1500, mean=400, sd=80, lo=50, hi=800, n_families=12, pool=500, seed=5)
m = B.compare_jaccard(B.SketchSet.from_host(h, off))
assert np.array_equal(m, orc.compare_all_pairs(h, off, nthreads=8)); print('stripe 1500x1500 identical')
" 2>&1 | tail -3
# 2. A/B on the 10 000-sketch matrix
for L in plain stripe stripe_upper cluster; do
  SMB_JOIN_LAYOUT=$L timeout 200 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_join_${L}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_join_${L}_${TAG}.json')); print('${L}: ms %.2f kernel_ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
for F in 0 1; do
  SMB_SKETCH_FUSED=$F timeout 300 python bench.py --workload sketch --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_sketch_fused${F}_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_sketch_fused${F}_${TAG}.json')); d=d.get('sketch', d); print('sketch fused=${F}: ms %.2f kernel_ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
for L in stripe stripe_upper; do
  SMB_JOIN_SORT=low32 SMB_JOIN_LAYOUT=$L timeout 200 python bench.py --workload compare --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_join_${L}_low32_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_join_${L}_low32_${TAG}.json')); print('${L} + low32 sort: ms %.2f kernel_ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
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
for L in stripe cluster; do
  SMB_JOIN_LAYOUT=$L ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum \
      --clock-control none -c 200 --csv --log-file gpurun_out/launches_${L}_${TAG}.csv \
      python bench.py --workload compare --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/launches_${L}_${TAG}.err
  tail -1 gpurun_out/launches_${L}_${TAG}.err
done
SMB_JOIN_LAYOUT=stripe ncu --set full --clock-control none --import-source on -k regex:join_stripe_kernel -c 1 \
    -o gpurun_out/stripe_${TAG} python bench.py --workload compare --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
# 4. range-partitioned search pass: correctness (every test that goes through one-vs-many), then A/B
echo "== tests with SMB_SEARCH_LAYOUT=ranges"
SMB_SEARCH_LAYOUT=ranges timeout 900 python -m pytest tests -q -m gpu -k "one_vs_many or search or gather or index or prefetch or zip" 2>&1 | tail -4
for L in plain ranges; do
  for W in search gather; do
    SMB_SEARCH_LAYOUT=$L timeout 400 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_${L}_${TAG}.json 2> /dev/null
    python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${L}_${TAG}.json')); print('${W} ${L}: %.2f ms'%d['ms_per_step'])"
  done
done
for W in search gather; do
  timeout 600 python bench.py --workload $W --index --steps 3 --warmup 3 > gpurun_out/bench_${W}_index_${TAG}.json 2> /dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_index_${TAG}.json')); print('${W} index: %.3f ms, build %.1f ms, %d distinct hashes'%(d['ms_per_step'], d['index']['build_ms'], d['index']['distinct_hashes']))"
done
SMB_SEARCH_LAYOUT=ranges ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_ranges_${TAG}.csv python bench.py --workload search --steps 1 --warmup 1 > /dev/null 2> gpurun_out/launches_ranges_${TAG}.err
ls -la gpurun_out | tail -12
