#!/bin/bash
# A/B of the PredictiveModel step-buffer reuse (AVC_PM_NO_WS=1: a fresh zero-filled arena per call) + the tests it touches
mkdir -p gpurun_out
for ws in 1 0; do
  if [[ $ws == 1 ]]; then export AVC_PM_NO_WS=1; else unset AVC_PM_NO_WS; fi
  echo "AVC_PM_NO_WS=${AVC_PM_NO_WS:-unset}"
  python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
  python bench.py --workload vsmask --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-260
done
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
