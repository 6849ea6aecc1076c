#!/bin/bash
mkdir -p gpurun_out
for a in "3 4" "3 2" "5 4" "148 4"; do echo "== grid/nsub $a"; timeout 60 python scripts/chain_debug2.py $a 2>&1 | grep -v Warn | tail -14; done > gpurun_out/r2m_debug.txt 2>&1
cat gpurun_out/r2m_debug.txt
timeout 300 python -m pytest tests/test_gpu_conv_chain.py -q 2>&1 | grep -E "RESULT|passed|failed|Error" | tail -40
