#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_sweep.sh N [config5]
# one distillation bench, one inference bench (and optionally config 5) on N GPUs of the box; JSON lines under gpurun_out/
N=$1; mkdir -p gpurun_out
tr() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@"; }
tr 29521 --steps 30 --warmup 10 --workload config3_distill --no-cpu-baseline > gpurun_out/sweep_distill_n$N.json 2> gpurun_out/sweep_distill_n$N.err; echo "distill rc=$?"
tr 29522 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/sweep_bench_n$N.json 2> gpurun_out/sweep_bench_n$N.err; echo "bench rc=$?"
if [ "$2" = "config5" ]; then tr 29523 --steps 10 --warmup 3 --workload config5_lidar --no-cpu-baseline > gpurun_out/sweep_config5_n$N.json 2> gpurun_out/sweep_config5_n$N.err; echo "config5 rc=$?"; fi
python - <<PY
import json
for f in ('sweep_distill_n$N', 'sweep_bench_n$N', 'sweep_config5_n$N'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step %.2f' % d['ms_per_step'], 'value %.3e' % d['value'], 'allreduce', d.get('allreduce'))
    except Exception as e:
        print(f, 'missing/failed', e)
PY
