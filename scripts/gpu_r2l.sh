#!/bin/bash
# Round 2, GPU call (4 GPUs): padded all-gather (>= 4 ranks) + 16-bit counters: verify and compare + sketch at N=4.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
bash scripts/gpu_multi.sh "4" ${1:-r2l} "both"
