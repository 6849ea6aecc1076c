#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv_chain.py -x -q > gpurun_out/r2g_chain.log 2>&1; rc=$?; echo "chain rc=$rc"
tail -3 gpurun_out/r2g_chain.log
if [ $rc -ne 0 ]; then grep -E "RESULT|Error|error" gpurun_out/r2g_chain.log | tail -8; exit 1; fi
rm -f gpurun_out/r2g_prof.txt
for k in "" "chain_dbg_skip=0x7"; do
  timeout 100 python scripts/conv_prof.py config2_200k 96 96 3 $k >> gpurun_out/r2g_prof.txt 2>&1
done
cat gpurun_out/r2g_prof.txt
timeout 150 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2g_knobs_96.txt 2>&1; cat gpurun_out/r2g_knobs_96.txt
OSB_CHAIN=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_chain.json 2> gpurun_out/r2g_bench_chain.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench_chain.json').read().strip().splitlines()[-1])
    print('chain ms/step', d['ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
