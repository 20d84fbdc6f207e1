#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_predictive_gpu.py -q -x 2>&1 | tail -30
AVC_PM_NO_PAIR=1 timeout 600 python -m pytest tests/test_predictive_gpu.py -q -x 2>&1 | tail -5
