#!/bin/bash
# round-2 validation: stress, full GPU suite, every bench line
mkdir -p gpurun_out
run() { echo "== $*"; timeout 80 python scripts/chain_debug3.py "${@}" 2>&1 | grep -v Warn | grep "^run" | cut -c1-150 | sort | uniq -c | sort -rn | head -3; }
{ run 3 32 32 3 9; run 5 64 64 3 9; run 3 128 128 3 9; run 7 96 96 3 9; } > gpurun_out/r2s_debug.txt 2>&1
cat gpurun_out/r2s_debug.txt
if grep -q "rows off vs fp64 [1-9]" gpurun_out/r2s_debug.txt; then echo "WRONG RESULTS"; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest_gpu.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err; echo "bench $name rc=$?"; tail -c 400 gpurun_out/r2s_bench_$name.json | head -c 400; echo; }
b default --steps 30 --warmup 5
b modules --steps 20 --warmup 5 --modules --no-cpu-baseline
b config1 --steps 20 --warmup 5 --workload config1_50k --no-cpu-baseline
b config4 --steps 20 --warmup 5 --workload config4_matterport --no-cpu-baseline
b config5 --steps 20 --warmup 5 --workload config5_lidar --no-cpu-baseline
b distill1 --steps 10 --warmup 3 --workload config3_distill --no-cpu-baseline
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2s_bench_')[1], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], d['unit'], 'e2e ms', d.get('e2e', {}).get('ms_per_step'),
              'conv ms', d.get('roofline', {}).get('kernel_ms_per_step'), 'l2 frac', d.get('roofline', {}).get('l2_lens', {}).get('frac'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
