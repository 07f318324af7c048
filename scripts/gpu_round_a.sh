#!/bin/bash
# one GPU call: full GPU test-suite, the widened-row measurements, ncu capture of the protein kernel
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r1n}
nproc
timeout 700 python -m pytest tests -q -m gpu --durations=12 -x 2>&1 | tail -25 | tee gpurun_out/tests_${TAG}.log
timeout 400 python scripts/bench_extra.py > gpurun_out/extra_${TAG}.json 2> gpurun_out/extra_${TAG}.err; tail -3 gpurun_out/extra_${TAG}.err; cat gpurun_out/extra_${TAG}.json
timeout 200 python scripts/ingest_bench.py --sigs 4000 --genomes 64 > gpurun_out/ingest_${TAG}.json 2> gpurun_out/ingest_${TAG}.err; cat gpurun_out/ingest_${TAG}.json; tail -2 gpurun_out/ingest_${TAG}.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hash_aa -s 6 -c 2 \
    -o gpurun_out/prof_aa_${TAG} -f python scripts/bench_extra.py --what protein > /dev/null 2> gpurun_out/prof_aa_${TAG}.err
tail -2 gpurun_out/prof_aa_${TAG}.err
ls -la gpurun_out/
