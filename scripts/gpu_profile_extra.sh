#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for W in search gather; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${W}.csv python bench.py --workload $W --steps 1 --warmup 3 \
    > gpurun_out/launches_${W}.json 2> gpurun_out/launches_${W}.err
done
ls -la gpurun_out | head -20
