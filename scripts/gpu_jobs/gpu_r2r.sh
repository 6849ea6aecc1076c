#!/bin/bash
mkdir -p gpurun_out
for a in "3 2" "5 2"; do echo "== grid/nsub $a"; timeout 60 python scripts/chain_debug2.py $a 2>&1 | grep -v Warn | tail -8; done > gpurun_out/r2r_debug.txt 2>&1
cat gpurun_out/r2r_debug.txt | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_conv_chain.py -q > gpurun_out/r2r_chain.log 2>&1; rc=$?; echo "chain rc=$rc"; grep -E "passed|failed" gpurun_out/r2r_chain.log | tail -2
if [ $rc -ne 0 ]; then grep -E "RESULT|Error" gpurun_out/r2r_chain.log | tail -20; exit 1; fi
rm -f gpurun_out/r2r_prof.txt
for k in "" "chain_dbg_skip=0x7"; do timeout 100 python scripts/conv_prof.py config2_200k 96 96 3 $k >> gpurun_out/r2r_prof.txt 2>&1; done
cat gpurun_out/r2r_prof.txt
timeout 150 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2r_knobs_96.txt 2>&1; cat gpurun_out/r2r_knobs_96.txt
OSB_CHAIN=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2r_layers_chain.txt 2>&1; head -3 gpurun_out/r2r_layers_chain.txt
OSB_CHAIN=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_chain.json 2> gpurun_out/r2r_bench_chain.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2r_bench_chain.json').read().strip().splitlines()[-1])
    print('chain ms/step', d['ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
