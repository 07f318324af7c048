#!/bin/bash
# Round 2, the remaining 4 GPU-minutes (1 GPU): counters at bank-spreading columns (stripe_col, carried by the tags) --
# A/B in one process against column order and against the independent algorithms (identical matrices asserted), then
# the stripe / full-size compare tests.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2p}
timeout 80 python tests/tools/ab_variants.py compare > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err; grep "^compare" gpurun_out/ab_${TAG}.err | cut -c1-200
timeout 75 python -m pytest tests/test_gpu_stripe.py tests/test_gpu_fullsize.py -q -m gpu -x -k "stripe or compare_10k or shards or take_rows" 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
