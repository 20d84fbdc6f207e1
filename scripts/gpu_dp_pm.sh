#!/bin/bash
# data-parallel PredictiveModel / VSMask trainer bench lines on the visible GPUs
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29541 --workload pm --steps 20 --warmup 3 > gpurun_out/bench_pm_${N}gpu.json 2> gpurun_out/bench_pm_${N}gpu.err; echo "pm rc=$?"; cut -c1-200 gpurun_out/bench_pm_${N}gpu.json
run 29542 --workload vsmask --steps 20 --warmup 3 > gpurun_out/bench_vsmask_${N}gpu.json 2> gpurun_out/bench_vsmask_${N}gpu.err; echo "vsmask rc=$?"; cut -c1-200 gpurun_out/bench_vsmask_${N}gpu.json
