#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25
for T in 512 1024 256; do for C in 512 128; do
  echo "== SMB_TILE_THREADS=$T SMB_TILE_COLS=$C"
  SMB_TILE_THREADS=$T SMB_TILE_COLS=$C timeout 300 python bench.py --workload compare --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],1), 'pairs/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
done; done
