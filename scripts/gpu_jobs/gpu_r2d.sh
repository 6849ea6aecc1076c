#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv_chain.py -x -q > gpurun_out/r2d_chain.log 2>&1; rc=$?; echo "chain rc=$rc"
tail -3 gpurun_out/r2d_chain.log
if [ $rc -ne 0 ]; then grep -E "RESULT|Error|error" gpurun_out/r2d_chain.log | tail -8; exit 1; fi
for k in "" "chain_dbg_skip=0x100" "chain_dbg_skip=0x7" "chain_dbg_skip=0x107"; do
  timeout 100 python scripts/conv_prof.py config2_200k 96 96 3 $k >> gpurun_out/r2d_prof.txt 2>&1
done
cat gpurun_out/r2d_prof.txt
timeout 150 python scripts/conv_knobs.py config2_200k 96 96 3 > gpurun_out/r2d_knobs_96.txt 2>&1; cat gpurun_out/r2d_knobs_96.txt
