#!/bin/bash
# Round 2, GPU call 5 (8 GPUs): the scaling run of the shipping paths.  verify at N=8, then compare + sketch at N=1,2,4,8
# and search / gather at N=1 and N=8.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
bash scripts/gpu_multi.sh "8" ${TAG} "both search gather"
bash scripts/gpu_multi.sh "4 2" ${TAG} "both"
bash scripts/gpu_multi.sh "1" ${TAG} "both search gather"
ls gpurun_out | tail -30
