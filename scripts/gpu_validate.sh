#!/bin/bash
# Round-end validation on the GPU box: full -m gpu suite, smoke, bench with the occupancy grid on / off, reference arm,
# launch list.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_validate.sh   (outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_grid.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/bench_grid_on.json
OSB_OCCGRID=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_grid_off.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref_grid.json
python - <<'PY'
import json
for nm in ("on","off"):
    try:
        d=json.load(open(f"gpurun_out/bench_grid_{nm}.json"))
        print(nm, "ms", round(d["ms_per_step"],3), d["step_ms_stats"]["device"], "e2e", round(d["e2e"]["value"]/1e6,1), "conv ms", round(d["roofline"]["kernel_ms_per_step"],2), "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(nm, "ERR", e, open(f"gpurun_out/bench_grid_{nm}.json").read()[-600:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 300 --csv --log-file gpurun_out/launches_grid.csv python bench.py --steps 1 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_bench_grid.log 2>&1
