#!/bin/bash
# Round 2, last GPU call (8 GPUs): compare + sketch at N=8 with padded all-gather + 16-bit counters (verify at N=8: r2i).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --workload both --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_both_n8_${TAG}.json 2> gpurun_out/bench_both_n8_${TAG}.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_both_n8_${TAG}.json") if l.startswith("{")][-1])
for x in (d, d.get("sketch", {})):
    print("N=8", x["metric"][:40], "value %.4g"%x["value"], "ms %.3f"%x["ms_per_step"], "kernel %.3f"%x["roofline"]["kernel_ms"], "e2e %.1f ms"%x["e2e"]["ms_per_step"], x.get("clocks"))
PY
