#!/bin/bash
# 8-GPU weak-scaling lines (one box): configs[3] emb 4096 x 80x512 sharded 512/GPU, and the default e2e replicas
mkdir -p gpurun_out
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload emb --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_emb_${N}gpu.json 2> gpurun_out/bench_emb_${N}gpu.err; echo "emb rc=$?"; cut -c1-300 gpurun_out/bench_emb_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_e2e_${N}gpu.json 2> gpurun_out/bench_e2e_${N}gpu.err; echo "e2e rc=$?"; cut -c1-300 gpurun_out/bench_e2e_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload fb --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fb_${N}gpu.json 2> gpurun_out/bench_fb_${N}gpu.err; echo "fb rc=$?"; cut -c1-300 gpurun_out/bench_fb_${N}gpu.json
