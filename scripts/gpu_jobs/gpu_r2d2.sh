#!/bin/bash
# distillation step on 2 GPUs (DDP over NCCL) and on 1 GPU with more steps; default bench on 2 GPUs
mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 10 --workload config3_distill --no-cpu-baseline > gpurun_out/r2d_distill_n1.json 2> gpurun_out/r2d_distill_n1.err; echo "n1 rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 10 --workload config3_distill --no-cpu-baseline > gpurun_out/r2d_distill_n2.json 2> gpurun_out/r2d_distill_n2.err; echo "n2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ('r2d_distill_n1', 'r2d_distill_n2', 'r2d_bench_n2'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step %.2f' % d['ms_per_step'], 'value %.3e' % d['value'], 'stats', {k: v for k, v in d['step_ms_stats'].items() if k != 'in_order'} if 'min' in d['step_ms_stats'] else '', 'allreduce', d.get('allreduce'))
        if 'in_order' in d['step_ms_stats']: print('   in order', d['step_ms_stats']['in_order'])
    except Exception as e:
        print(f, 'failed', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
