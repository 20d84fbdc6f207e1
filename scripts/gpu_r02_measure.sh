#!/bin/bash
# Round-2 measurement pass on one GPU: bench (both arms), pm / vsmask lines, ncu launch list of the bench command,
# ncu --set full captures of the new TMA + tcgen05 kernels (PredictiveModel step) and of the dominant conv_tc launches.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py --steps 1500 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
for w in pm vsmask; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_${w}_1gpu.json 2> gpurun_out/bench_${w}_1gpu.err; echo "$w rc=$?"; cut -c1-250 gpurun_out/bench_${w}_1gpu.json
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
TGT="python scripts/pm_target.py 256"
timeout 200 $TGT > gpurun_out/ncu_pm_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv2d_tc_kernel|wgrad_tc_kernel" -s 58 -c 12 -o gpurun_out/r02f_pm_tc_kernels -f $TGT > gpurun_out/ncu_pm_tc.log 2>&1
echo "ncu pm full rc=$?"
TGT2="python scripts/ncu_target.py emb 128 512 2"
timeout 200 $TGT2 > gpurun_out/ncu_tc_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 41 -c 6 -o gpurun_out/r02f_conv_tc_emb_b128_top -f $TGT2 > gpurun_out/ncu_tc.log 2>&1
echo "ncu conv_tc full rc=$?"; ls -la gpurun_out/*.ncu-rep
echo "total $(( $(date +%s)-t0 )) s"
