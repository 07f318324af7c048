#!/bin/bash
# ncu evidence for the two dominant kernels (B200_PROFILING.md recipe); outputs under gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1}
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.bench.json 2> gpurun_out/launches_${TAG}.err
# 2. full capture of the intersection kernels (join count kernel; tile kernel with SMB_COMPARE_ALGO=tile)
#    and the hash kernels (3 launches: k=21,31,51)
ncu --set full --clock-control none --import-source on -k regex:join_count -s 2 -c 1 \
    -o gpurun_out/prof_join_${TAG} -f python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof_join_${TAG}.err
SMB_COMPARE_ALGO=tile ncu --set full --clock-control none --import-source on -k regex:pairwise_tile -s 2 -c 1 \
    -o gpurun_out/prof_tile_${TAG} -f python bench.py --workload compare --steps 1 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof_tile_${TAG}.err
ncu --set full --clock-control none --import-source on -k regex:hash_kmers -s 3 -c 3 \
    -o gpurun_out/prof_hash_${TAG} -f python bench.py --workload sketch --steps 1 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof_hash_${TAG}.err
ls -la gpurun_out/
for f in gpurun_out/*_${TAG}.err; do tail -n 3 "$f"; done
