#!/bin/bash
# round-2 call A: correctness of the persistent chain kernel, baseline-size parity, A/B bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_conv_chain.py -x -q -s > gpurun_out/r2a_chain.log 2>&1; echo "chain rc=$?" >> gpurun_out/r2a_rc.txt
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_baseline_sizes.py -x -q -s > gpurun_out/r2a_engine.log 2>&1; echo "engine rc=$?" >> gpurun_out/r2a_rc.txt
OSB_CHAIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_legacy.json 2> gpurun_out/r2a_bench_legacy.err; echo "bench0 rc=$?" >> gpurun_out/r2a_rc.txt
OSB_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_chain.json 2> gpurun_out/r2a_bench_chain.err; echo "bench1 rc=$?" >> gpurun_out/r2a_rc.txt
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=400 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_chain400.json 2> gpurun_out/r2a_bench_chain400.err; echo "bench2 rc=$?" >> gpurun_out/r2a_rc.txt
OSB_CHAIN=1 OSB_CHAIN_MAX_TILES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_chain0.json 2> gpurun_out/r2a_bench_chain0.err; echo "bench3 rc=$?" >> gpurun_out/r2a_rc.txt
cat gpurun_out/r2a_rc.txt
tail -5 gpurun_out/r2a_chain.log
