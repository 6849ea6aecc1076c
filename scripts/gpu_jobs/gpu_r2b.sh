#!/bin/bash
mkdir -p gpurun_out
OSB_CHAIN=0 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_legacy.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_nsub=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_nsub1.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_dbg_skip=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_noA.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_dbg_skip=3 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_noAB.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_dbg_skip=7 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_noABM.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_dbg_skip=8 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_nostore.txt 2>&1
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 KNOBS=chain_sa=4 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain0_sa4.txt 2>&1
OSB_CHAIN=1 timeout 200 python scripts/layer_times.py > gpurun_out/r2b_layers_chain.txt 2>&1
head -3 gpurun_out/r2b_layers_*.txt
timeout 300 python -m pytest tests/test_gpu_fast_eval.py -x -q -s > gpurun_out/r2b_fast_eval.log 2>&1; echo "fast_eval rc=$?"
tail -15 gpurun_out/r2b_fast_eval.log
