#!/bin/bash
# one PredictiveModel iteration on the GPU: parity tests, the two bench lines, a per-launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py -q -x 2>&1 | tail -3
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
python bench.py --workload vsmask --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/pm_launches.csv python scripts/pm_target.py 256 > gpurun_out/ncu_pm_list.log 2>&1; echo "ncu rc=$?"
