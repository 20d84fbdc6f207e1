#!/bin/bash
# A/B of the TMA-fed in-conv dgrad (conv2d_tc) against conv_tc's nine column passes, then the attack parity tests
for m in on off; do
  if [[ $m == off ]]; then export AVC_NO_TMA_INCONV=1; else unset AVC_NO_TMA_INCONV; fi
  echo "== TMA in-conv dgrad $m"
  for c in "emb 128 512" "emb 512 512" "fb 64 256"; do timeout 200 python scripts/batched_probe.py $c | head -2; done
done
unset AVC_NO_TMA_INCONV
timeout 200 python scripts/batched_probe.py emb 128 512 v | sed -n 40,48p
timeout 900 python -m pytest tests/test_attacks_gpu.py tests/test_models_gpu.py tests/test_vsmask_train_gpu.py -q -x 2>&1 | tail -2
