#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_pytest_gpu.log
timeout 300 python scripts/distill_profile.py > gpurun_out/r3a_distill_profile.txt 2>&1; head -40 gpurun_out/r3a_distill_profile.txt | cut -c1-170
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r3a_bench_$name.json 2> gpurun_out/r3a_bench_$name.err; echo "bench $name rc=$?"; }
b default --steps 30 --warmup 5
b config4 --steps 20 --warmup 5 --workload config4_matterport --no-cpu-baseline
b config5 --steps 20 --warmup 5 --workload config5_lidar --no-cpu-baseline
timeout 200 python scripts/bench_next_rows.py container 2>/dev/null | cut -c1-600
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r3a_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r3a_bench_')[1], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e ms', d.get('e2e', {}).get('ms_per_step'),
              'conv ms', d.get('roofline', {}).get('kernel_ms_per_step'), 'points ms', (d.get('e2e_points') or {}).get('ms_per_step'), 'folded', (d.get('extra') or {}).get('folded_head_ms_per_step'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
