#!/bin/bash
# One development iteration on the GPU box: parity tests, batched probe (product build), issuer profile (instrumented build).
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s)-t0 )) s)"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/batched_probe.py emb 128 512 v > gpurun_out/probe_emb.log 2>&1; echo "probe emb rc=$?"; head -8 gpurun_out/probe_emb.log
timeout 300 python scripts/batched_probe.py fb 64 256 v > gpurun_out/probe_fb.log 2>&1; echo "probe fb rc=$?"; head -8 gpurun_out/probe_fb.log
if [[ -f attack_vc_b200/libavc_b200_prof.so ]]; then
  AVC_LIB=attack_vc_b200/libavc_b200_prof.so AVC_TC_DBG=32 timeout 200 python scripts/batched_probe.py emb 128 512 v > gpurun_out/tcprof_32.log 2>&1; echo "prof rc=$?"
  grep "conv_tc cta0" gpurun_out/tcprof_32.log | tail -41 | head -16
fi
if [[ -n "$BENCH" ]]; then
  timeout 600 python bench.py --steps 1500 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json
fi
echo "total $(( $(date +%s)-t0 )) s"
