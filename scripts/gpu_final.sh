#!/bin/bash
# last pass of a round: parity tests, smoke, bench (both arms), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 1500 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-160 gpurun_out/bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
