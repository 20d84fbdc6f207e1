#!/bin/bash
# PredictiveModel: parity tests, then the bench line with the default 3xTF32 convs and with the opt-in bf16 pair planes (AVC_PM_PAIR=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py tests/test_audio_gpu.py -q -x 2>&1 | tail -3
for np in 0 1; do
  if [[ $np == 1 ]]; then export AVC_PM_PAIR=1; else unset AVC_PM_PAIR; fi
  echo "AVC_PM_PAIR=${AVC_PM_PAIR:-unset}"
  python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
done
unset AVC_PM_PAIR
python bench.py --workload vsmask --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-260
