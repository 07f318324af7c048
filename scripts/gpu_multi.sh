#!/bin/bash
# N-GPU bench (torchrun, NCCL) -- run with: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "rc=$?"
tail -15 gpurun_out/bench_n${N}.err
python - <<PY
import json
try:
    lines=[l for l in open("gpurun_out/bench_n${N}.json") if l.startswith("{")]
    print("stdout lines:", sum(1 for _ in open("gpurun_out/bench_n${N}.json")))
    d=json.loads(lines[-1])
    for x in (d, d.get("sketch", {})):
        if x: print(x["metric"], "n_gpus", x["n_gpus"], "value %.4g"%x["value"], "ms %.2f"%x["ms_per_step"], "e2e %.4g (%.1f ms)"%(x["e2e"]["value"], x["e2e"]["ms_per_step"]), "kernel_ms %.2f"%x["roofline"]["kernel_ms"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_n${N}.json").read()[:2000])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tests/tools/multi_gpu_verify.py > gpurun_out/verify_n${N}.log 2>&1
grep -E "multi-GPU verify|Error|assert" gpurun_out/verify_n${N}.log | head -10
# experimental: stripe layout, every rank counts its own block of rows (no all-reduce); verify, then bench
SMB_JOIN_LAYOUT=stripe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    tests/tools/multi_gpu_verify.py > gpurun_out/verify_stripe_n${N}.log 2>&1
grep -E "multi-GPU verify|Error|assert" gpurun_out/verify_stripe_n${N}.log | head -5
SMB_JOIN_LAYOUT=stripe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --workload compare --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stripe_n${N}.json 2> gpurun_out/bench_stripe_n${N}.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_stripe_n${N}.json') if l.startswith('{')][-1]); print('stripe n=${N}: ms %.2f e2e %.1f ms'%(d['ms_per_step'], d['e2e']['ms_per_step']))"
# configs[3] / configs[4] sharded by subject over the N GPUs (ShardedDatabase), without and with the inverted index
for W in search gather; do
  for IDX in "" "--index"; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
        bench.py --gpus $N --workload $W $IDX --steps 3 --warmup 3 > gpurun_out/bench_${W}${IDX}_n${N}.json 2> gpurun_out/bench_${W}${IDX}_n${N}.err
    python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${W}${IDX}_n${N}.json') if l.startswith('{')][-1]); print('${W} ${IDX} n=${N}: %.2f ms'%d['ms_per_step'], d.get('index',''))"
  done
done
