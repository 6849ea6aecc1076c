#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; timeout 80 python scripts/chain_debug3.py "${@}" 2>&1 | grep -v Warn | grep "^run" | cut -c1-150 | sort | uniq -c | sort -rn | head -4; }
{
run 3 32 32 3 12
run 5 64 64 3 12
run 3 128 128 3 12
run 7 96 96 3 12
run 3 32 32 3 12 chain_pipes=1 chain_dbg_skip=0x400
} > gpurun_out/r2w_debug.txt 2>&1
cat gpurun_out/r2w_debug.txt
timeout 300 python -m pytest tests/test_gpu_conv_chain.py -q -x > gpurun_out/r2t_chain.log 2>&1; rc=$?; echo "chain rc=$rc"; grep -E "passed|failed" gpurun_out/r2t_chain.log | tail -2
if [ $rc -ne 0 ]; then grep -E "RESULT|Error|assert" gpurun_out/r2t_chain.log | tail -20; exit 1; fi
rm -f gpurun_out/r2t_prof.txt
for k in "" "chain_dbg_skip=0x80"; do timeout 100 python scripts/conv_prof.py config2_200k 96 96 3 $k >> gpurun_out/r2t_prof.txt 2>&1; done
cat gpurun_out/r2t_prof.txt
timeout 200 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2t_knobs_96.txt 2>&1; cat gpurun_out/r2t_knobs_96.txt
OSB_CHAIN=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2t_layers_chain.txt 2>&1; head -3 gpurun_out/r2t_layers_chain.txt
