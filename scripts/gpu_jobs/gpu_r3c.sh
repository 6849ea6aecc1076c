#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -x > gpurun_out/r3c_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r3c_tests.log
if [ $rc -ne 0 ]; then grep -E "Error|assert" gpurun_out/r3c_tests.log | tail -20; exit 1; fi
timeout 300 python scripts/distill_profile.py > gpurun_out/r3c_distill_profile.txt 2>&1; head -22 gpurun_out/r3c_distill_profile.txt | grep -v Warn | cut -c1-170
timeout 400 python bench.py --steps 30 --warmup 10 --workload config3_distill --no-cpu-baseline > gpurun_out/r3c_distill_n1.json 2> gpurun_out/r3c_distill_n1.err; echo "distill rc=$?"
python - <<'PY'
import json
for f in ('r3c_distill_n1',):
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    st = d['step_ms_stats']
    print(f, 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], {k: v for k, v in st.items() if k != 'in_order'})
PY
