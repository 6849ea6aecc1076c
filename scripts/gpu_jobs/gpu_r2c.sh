#!/bin/bash
# correctness gate first; timing only if it passes.  Short timeouts: a trapped kernel costs seconds, not minutes.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv_chain.py -x -q > gpurun_out/r2c_chain.log 2>&1; rc=$?; echo "chain rc=$rc"
tail -3 gpurun_out/r2c_chain.log
if [ $rc -ne 0 ]; then grep RESULT gpurun_out/r2c_chain.log | tail -5; exit 1; fi
timeout 150 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2c_knobs_96.txt 2>&1
timeout 150 python scripts/conv_knobs.py config2_200k 32 32 3 > gpurun_out/r2c_knobs_32.txt 2>&1
cat gpurun_out/r2c_knobs_96.txt gpurun_out/r2c_knobs_32.txt
OSB_CHAIN=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_chain.json 2> gpurun_out/r2c_bench_chain.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_chain.json').read().strip().splitlines()[-1])
    print('chain ms/step', d['ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
timeout 200 python -m pytest tests/test_gpu_fast_eval.py tests/test_gpu_wgrad_tc.py -q > gpurun_out/r2c_fast_eval.log 2>&1; echo "fast_eval+wgrad rc=$?"; tail -25 gpurun_out/r2c_fast_eval.log
