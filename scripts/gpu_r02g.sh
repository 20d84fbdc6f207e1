#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_g.json'))
print(d["value"], {k:(v.get("frac_of_hbm_peak") or v.get("ms_per_step")) for k,v in d["batched"].items()})
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_small -s 120 -c 8 -o gpurun_out/r02g_conv_small_e2e_b1 -f $CMD > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
TGT2="python scripts/ncu_target.py emb 128 512 2"
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --clock-control none -k regex:"conv_tc" -c 200 -o gpurun_out/r02g_conv_tc_emb_b128_all -f $TGT2 > gpurun_out/ncu_tc_all.log 2>&1; echo "ncu tc all rc=$?"
TGT="python scripts/pm_target.py 256"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 10 -c 10 -o gpurun_out/r02g_pm_wgrad_tc -f $TGT > gpurun_out/ncu_pm_wg.log 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
