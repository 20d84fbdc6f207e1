#!/bin/bash
# final check of the committed tree: full GPU suite, smoke, the default bench line, the PredictiveModel / trainer lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_final.json
python bench.py --workload pm --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pm_final.json; cut -c1-200 gpurun_out/bench_pm_final.json
python bench.py --workload vsmask --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vsmask_final.json; cut -c1-220 gpurun_out/bench_vsmask_final.json
