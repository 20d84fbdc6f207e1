#!/bin/bash
# N-GPU bench lines of the sharded workloads (N = number of visible GPUs): PredictiveModel (data parallel: sync BatchNorm + gradient
# all-reduce), VSMask trainer, emb attack (cfg4, sharded leg), plus the NCCL parity check.
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 600 python -m pytest tests/test_nccl_gpu.py -x -q -s > gpurun_out/pytest_nccl_${N}gpu.log 2>&1; echo "nccl rc=$?"; grep -E "PASS|FAIL|passed|failed" gpurun_out/pytest_nccl_${N}gpu.log
run 29541 --workload pm --steps 20 --warmup 3 > gpurun_out/bench_pm_${N}gpu.json 2> gpurun_out/bench_pm_${N}gpu.err; echo "pm rc=$?"; cut -c1-200 gpurun_out/bench_pm_${N}gpu.json
run 29542 --workload vsmask --steps 20 --warmup 3 > gpurun_out/bench_vsmask_${N}gpu.json 2> gpurun_out/bench_vsmask_${N}gpu.err; echo "vsmask rc=$?"; cut -c1-200 gpurun_out/bench_vsmask_${N}gpu.json
run 29543 --workload emb --steps 30 --warmup 5 > gpurun_out/bench_emb_${N}gpu.json 2> gpurun_out/bench_emb_${N}gpu.err; echo "emb rc=$?"; cut -c1-200 gpurun_out/bench_emb_${N}gpu.json
run 29544 --workload fb --steps 30 --warmup 5 > gpurun_out/bench_fb_${N}gpu.json 2> gpurun_out/bench_fb_${N}gpu.err; echo "fb rc=$?"; cut -c1-200 gpurun_out/bench_fb_${N}gpu.json
