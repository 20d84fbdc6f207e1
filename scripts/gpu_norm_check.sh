#!/bin/bash
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n.json'))
print("value", d["value"], {k:round(v.get("frac_of_hbm_peak"),4) for k,v in d["batched"].items() if "norm" in k}, "fb norm", d["batched"]["fb_B64_T256"].get("norm_frac_of_hbm_peak"), "fb ms", d["batched"]["fb_B64_T256"]["ms_per_step"])
PY
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_attacks_gpu.py tests/test_models_gpu.py -q -x 2>&1 | tail -2
