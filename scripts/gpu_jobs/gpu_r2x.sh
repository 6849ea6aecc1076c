#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; timeout 80 python scripts/chain_debug3.py "${@}" 2>&1 | grep -v Warn | grep "^run" | cut -c1-150 | sort | uniq -c | sort -rn | head -4; }
{
run 3 32 32 3 9
run 5 64 64 3 9
run 7 96 96 3 9
} > gpurun_out/r2x_debug.txt 2>&1
cat gpurun_out/r2x_debug.txt
if grep -q "rows off vs fp64 [1-9]" gpurun_out/r2x_debug.txt; then echo "WRONG RESULTS"; exit 1; fi
rm -f gpurun_out/r2x_prof.txt
for k in ""; do timeout 100 python scripts/conv_prof.py config2_200k 96 96 3 $k >> gpurun_out/r2x_prof.txt 2>&1; done
cat gpurun_out/r2x_prof.txt
timeout 200 python scripts/conv_knobs.py config2_200k 96 96 3 2>&1 | grep -v "^Traceback\|^  File\|^    \|Warn" > gpurun_out/r2x_knobs_96.txt; head -16 gpurun_out/r2x_knobs_96.txt
timeout 300 python -m pytest tests/test_gpu_conv_chain.py -q -x > gpurun_out/r2x_chain.log 2>&1; rc=$?; echo "chain rc=$rc"; grep -E "passed|failed" gpurun_out/r2x_chain.log | tail -2
