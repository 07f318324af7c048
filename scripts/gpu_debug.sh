#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_minhash_port.py -x -q -m gpu 2>&1 | tail -40
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -8
