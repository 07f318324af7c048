#!/bin/bash
# compute-sanitizer (memcheck) over the small-input GPU tests
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_minhash_port.py -q -m gpu -x \
    -k "not ecoli and not three_k and not large" 2>&1 | tail -5
echo "rc=$?"
grep -c "ERROR SUMMARY" gpurun_out/memcheck.log; grep "ERROR SUMMARY" gpurun_out/memcheck.log | sort | uniq -c | head; grep -m5 -A12 "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck.log | head -60
