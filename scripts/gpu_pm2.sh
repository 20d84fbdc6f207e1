#!/bin/bash
# 2-GPU pass: NCCL parity check (sharded attacks, header, data-parallel trainer), pm / vsmask bench lines at N = 2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nccl_gpu.py -x -q -s > gpurun_out/pytest_nccl.log 2>&1; echo "nccl rc=$?"; grep -E "PASS|FAIL|passed|failed" gpurun_out/pytest_nccl.log
for w in pm vsmask; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_1gpu.json 2> gpurun_out/bench_${w}_1gpu.err; echo "$w 1gpu rc=$?"; cut -c1-330 gpurun_out/bench_${w}_1gpu.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_2gpu.json 2> gpurun_out/bench_${w}_2gpu.err; echo "$w 2gpu rc=$?"; cut -c1-330 gpurun_out/bench_${w}_2gpu.json; tail -3 gpurun_out/bench_${w}_2gpu.err
done
