#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/hbm_probe.py | tee gpurun_out/hbm_probe.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "norm or adam or update" 2>&1 | tail -2
