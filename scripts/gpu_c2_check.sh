#!/bin/bash
timeout 900 python -m pytest tests/test_predictive_gpu.py tests/test_vsmask_train_gpu.py tests/test_attacks_gpu.py -q -x 2>&1 | tail -2
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-230
python bench.py --workload vsmask --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-260
for c in "emb 128 512" "emb 512 512"; do timeout 200 python scripts/batched_probe.py $c | head -2; done
timeout 200 python scripts/batched_probe.py emb 128 512 v | sed -n 44,46p
if [[ -f attack_vc_b200/libavc_b200_prof.so ]]; then
  AVC_LIB=attack_vc_b200/libavc_b200_prof.so timeout 200 python scripts/batched_probe.py emb 128 512 > gpurun_out/c2prof_emb.log 2>&1; grep "conv2d_tc cta0" gpurun_out/c2prof_emb.log | tail -3
  AVC_LIB=attack_vc_b200/libavc_b200_prof.so timeout 200 python scripts/pm_target.py 256 > gpurun_out/c2prof_pm.log 2>&1; grep "conv2d_tc cta0" gpurun_out/c2prof_pm.log | tail -162 | head -30
fi
