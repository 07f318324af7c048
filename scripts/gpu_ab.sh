#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
for V in split u64; do
  echo "== SMB_TILE_VARIANT=$V"
  SMB_TILE_VARIANT=$V timeout 300 python bench.py --workload compare --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],1), 'pairs/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
done
