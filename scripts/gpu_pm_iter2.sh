#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py -q -x 2>&1 | tail -3
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/pm_launches.csv python scripts/pm_target.py 256 > gpurun_out/ncu_pm_list.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pm_conv_kernel|pm_convT_co1" -s 3 -c 3 -o gpurun_out/r02n_pm_c1 -f python scripts/pm_target.py 256 > gpurun_out/ncu_pm_c1.log 2>&1; echo "ncu2 rc=$?"
