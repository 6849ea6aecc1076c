#!/bin/bash
# round-2 profiles: smoke, reference arm, launch list of a bench step, ncu --set full of the persistent kernel on the level-0 96->96 layer
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2n_bench_reference.json; tail -c 300 gpurun_out/r2n_bench_reference.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu_bench.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2n_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:k_conv_chain -s 2 -c 1 -o gpurun_out/r2n_full_chain_96 python scripts/prof_chain_one.py 96 96 3 > gpurun_out/r2n_ncu_full.log 2>&1; echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
