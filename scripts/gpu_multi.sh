#!/bin/bash
# N-GPU verification + bench (torchrun, NCCL) -- run with: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [tag] [workloads]'
# For every N in the list (default: just N) : multi_gpu_verify.py (results bit-identical to the oracle / the single-GPU
# result), then bench.py for the default line (compare + sketch) and the search / gather workloads.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
NS=${1:-2}; TAG=${2:-r2}; WL=${3:-"both search gather"}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
PORT=29511
for N in $NS; do
  if [ "$N" -gt 1 ]; then
    RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
    PORT=$((PORT+1))
    timeout 600 $RUN $PORT tests/tools/multi_gpu_verify.py > gpurun_out/verify_n${N}_${TAG}.log 2>&1
    grep -E "multi-GPU verify|Error|assert|Traceback" gpurun_out/verify_n${N}_${TAG}.log | head -6
  else
    RUN=""
  fi
  for W in $WL; do
    PORT=$((PORT+1))
    if [ "$N" -gt 1 ]; then CMD="$RUN $PORT bench.py"; else CMD="python bench.py"; fi
    timeout 900 $CMD --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_n${N}_${TAG}.json 2> gpurun_out/bench_${W}_n${N}_${TAG}.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${W}_n${N}_${TAG}.json") if l.startswith("{")][-1])
    for x in (d, d.get("sketch", {})):
        if x: print("N=$N", x["metric"][:40], "value %.4g"%x["value"], "ms %.3f"%x["ms_per_step"], "e2e %.1f ms"%x["e2e"]["ms_per_step"], "clocks", (x.get("clocks") or {}).get("samples"))
except Exception as e:
    print("N=$N $W: no json:", e); print(open("gpurun_out/bench_${W}_n${N}_${TAG}.err").read()[-1500:])
PY
  done
done
