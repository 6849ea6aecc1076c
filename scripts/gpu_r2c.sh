#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_chain.py -x -q > gpurun_out/r2c_chain.log 2>&1; echo "chain rc=$?"
tail -3 gpurun_out/r2c_chain.log
timeout 300 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2c_knobs_96.txt 2>&1
timeout 300 python scripts/conv_knobs.py config2_200k 32 32 3 > gpurun_out/r2c_knobs_32.txt 2>&1
cat gpurun_out/r2c_knobs_96.txt gpurun_out/r2c_knobs_32.txt
OSB_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_chain.json 2> gpurun_out/r2c_bench_chain.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_chain.json').read().strip().splitlines()[-1])
print('chain ms/step', d['ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'])
PY
timeout 300 python -m pytest tests/test_gpu_fast_eval.py -x -q > gpurun_out/r2c_fast_eval.log 2>&1; echo "fast_eval rc=$?"; tail -5 gpurun_out/r2c_fast_eval.log
