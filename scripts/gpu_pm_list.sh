#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pm_target.py 256 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/pm_launches.csv python scripts/pm_target.py 256 > gpurun_out/ncu_pm_list.log 2>&1; echo "ncu rc=$?"
