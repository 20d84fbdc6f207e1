#!/bin/bash
# Measurement pass on the GPU box (one gpurun call): bench (both arms), ncu launch list of the bench command,
# ncu --set full of a few conv_tc launches of a batched emb attack.  Every ncu run follows a plain run that exited 0.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py --steps 1500 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
TGT="python scripts/ncu_target.py emb 128 512 2"
timeout 200 $TGT > gpurun_out/ncu_tc_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 41 -c 8 -o gpurun_out/conv_tc_emb_b128 -f $TGT > gpurun_out/ncu_tc.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
echo "total $(( $(date +%s)-t0 )) s"
