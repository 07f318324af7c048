#!/bin/bash
# optional workloads: search (configs[3]) and gather (configs[4]) + the whole GPU test-suite
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
for W in search gather; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  tail -2 gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json
done
