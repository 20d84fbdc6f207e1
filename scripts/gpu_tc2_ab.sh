#!/bin/bash
# A/B of the two tcgen05 conv kernels on the batched workloads (auto selection, N=128 kernel only, N=256 kernel wherever legal).
mkdir -p gpurun_out
for mode in auto 1 2; do
  echo "== AVC_TC2=$mode"
  if [[ $mode == auto ]]; then unset AVC_TC2; else export AVC_TC2=$mode; fi
  for c in "emb 128 512" "emb 512 512" "fb 64 256" "e2e 32 256"; do timeout 200 python scripts/batched_probe.py $c | head -2; done
done
