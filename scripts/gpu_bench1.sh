#!/bin/bash
# smoke + default bench on one B200; outputs under gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1000 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -8 gpurun_out/bench1.err
cat gpurun_out/bench1.json
