#!/bin/bash
# The other BASELINE configurations on the final build (not bench lines: evidence for DESIGN.md).
mkdir -p gpurun_out
timeout 150 python bench.py --workload config4_matterport --k-text 160 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config4.json
timeout 200 python bench.py --workload config5_lidar --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config5.json
timeout 100 python bench.py --workload config1_50k --arch MinkUNet18A --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config1.json
python - <<'PY'
import json
for c in ("config4","config5","config1"):
    try:
        d=json.load(open(f"gpurun_out/bench_{c}.json"))
        print(c, d["config"]["workload"][:60], "ms", round(d["ms_per_step"],3), "Mvox/s", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), "TF", round(d["roofline"]["tflops"],1), "frac", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(c, "ERR", e, open(f"gpurun_out/bench_{c}.json").read()[-400:])
PY
