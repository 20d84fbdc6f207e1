#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py tests/test_kernels_gpu.py -q -x -k "not tc_ or wgrad" 2>&1 | tail -2
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/pm_launches.csv python scripts/pm_target.py 256 > gpurun_out/ncu_pm_list.log 2>&1; echo "ncu rc=$?"
