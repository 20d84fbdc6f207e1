#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
AVC_PM_PAIR=1 python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
AVC_PM_PAIR=1 timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py -q 2>&1 | tail -12
