#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): the scaling run of the committed tree: verify at N=8, compare + sketch at N=8, 4, 2,
# search / gather at N=8, gather at N=1 (N=1 of the other workloads: r2h).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2i}
bash scripts/gpu_multi.sh "8" ${TAG} "both search gather"
bash scripts/gpu_multi.sh "4 2" ${TAG} "both"
bash scripts/gpu_multi.sh "1" ${TAG} "gather both"
ls gpurun_out | grep ${TAG} | wc -l
