#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv_chain.py tests/test_gpu_engine.py tests/test_gpu_fast_eval.py -q -x > gpurun_out/r2y_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2y_tests.log
if [ $rc -ne 0 ]; then grep -E "RESULT|Error|assert" gpurun_out/r2y_tests.log | tail -20; exit 1; fi
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 timeout 200 python scripts/layer_times.py > gpurun_out/r2y_layers_chain0.txt 2>&1; head -2 gpurun_out/r2y_layers_chain0.txt | cut -c1-200
OSB_CHAIN=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2y_layers_chain.txt 2>&1; head -2 gpurun_out/r2y_layers_chain.txt | cut -c1-250
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['step_ms_stats']['device'])
PY
