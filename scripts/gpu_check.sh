#!/bin/bash
# Run on the GPU box via gpurun: parity tests, smoke, bench (both arms), ncu launch list.
# Usage: scripts/gpu_check.sh [tests|bench|ncu|all]
mode=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
if [[ $mode == all || $mode == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
fi
if [[ $mode == all || $mode == bench ]]; then
  timeout 600 python bench.py --steps ${STEPS:-1500} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
fi
if [[ $mode == all || $mode == ncu ]]; then
  CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
  timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
  echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
fi
