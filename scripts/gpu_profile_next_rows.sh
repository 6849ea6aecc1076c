#!/bin/bash
# ncu --set full captures of the kernels added late in round 1 (fusion gather, grid stem, grid kernel map).
# Usage: gpurun --timeout 900 -- bash scripts/gpu_profile_next_rows.sh     (reports under gpurun_out/)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 250 $NCU -k regex:k_fusion_gather -s 4 -c 1 -o gpurun_out/prof_fusion_gather python scripts/bench_next_rows.py > gpurun_out/prof_fusion.log 2>&1
timeout 250 $NCU -k regex:k_conv_stem -s 6 -c 1 -o gpurun_out/prof_stem_grid python bench.py --steps 1 --warmup 6 --no-cpu-baseline > gpurun_out/prof_stem.log 2>&1
timeout 250 $NCU -k regex:k_kernel_map_grid -s 55 -c 2 -o gpurun_out/prof_kmap_grid python bench.py --steps 1 --warmup 6 --no-cpu-baseline > gpurun_out/prof_kmap.log 2>&1
ls -la gpurun_out/*.ncu-rep
