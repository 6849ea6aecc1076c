#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_container.py -q > gpurun_out/r2z_container.log 2>&1; echo "container tests rc=$?"; tail -2 gpurun_out/r2z_container.log
timeout 600 python scripts/bench_next_rows.py voxelize remap container > gpurun_out/r2z_next_rows.jsonl 2> gpurun_out/r2z_next_rows.err; echo "next rows rc=$?"; cut -c1-700 gpurun_out/r2z_next_rows.jsonl; tail -3 gpurun_out/r2z_next_rows.err
for fs in 2 3; do OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_force_split=$fs timeout 200 python scripts/layer_times.py > gpurun_out/r2z_layers_split$fs.txt 2>&1; done
paste <(sed -n 2,63p gpurun_out/r2y_layers_chain0.txt | cut -c1-50) <(sed -n 2,63p gpurun_out/r2z_layers_split2.txt | cut -c38-50) <(sed -n 2,63p gpurun_out/r2z_layers_split3.txt | cut -c38-50) | sed -n 1,5p
paste <(sed -n 2,63p gpurun_out/r2y_layers_chain0.txt | cut -c1-50) <(sed -n 2,63p gpurun_out/r2z_layers_split2.txt | cut -c38-50) <(sed -n 2,63p gpurun_out/r2z_layers_split3.txt | cut -c38-50) | sed -n 50,62p
