"""Time ONE convolution shape through the persistent kernel under a list of tuning-knob settings.
    python scripts/conv_knobs.py [n_rows_workload] cin cout K -- prints us per launch (median of 20) per setting."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import synth, tc  # noqa: E402
from openscene_b200.coords import CoordinateManager  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'config2_200k'
cin, cout, ks = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (96, 96, 3)))
dev = torch.device('cuda:0')
c = torch.from_numpy(synth.scene(workload)).to(dev)
cm = CoordinateManager(c)
n = cm.sets[1].n
nbr = cm.kernel_map(1, 1, ks).nbr if ks > 1 else None
g = torch.Generator(device=dev).manual_seed(0)
x = tc.to_split(torch.randn(n, cin, device=dev, generator=g))
w = torch.randn(ks ** 3, cin, cout, device=dev, generator=g) * 0.05
wt, wp = tc.pack_weight_tiles(w), tc.pack_weights(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


print(f'# {workload} n={n} {cin}->{cout} k={ks}')
print(f'legacy k_conv_tc                         {timed(lambda: tc.conv_tc(x, cin, None, 0, nbr, n, ks ** 3, wp, cout)):8.1f} us')
SETS = [
    ('chain default (noinc arrivals, consumer-side fence)', {}),

    ('nsub=1', {'chain_nsub': 1}),
    ('no A,B,MMA (0x7)', {'chain_dbg_skip': 0x7}),
    ('no A,B (0x3)', {'chain_dbg_skip': 0x3}),
    ('no A (0x1)', {'chain_dbg_skip': 0x1}),
    ('no B (0x2)', {'chain_dbg_skip': 0x2}),
    ('no MMA (0x4)', {'chain_dbg_skip': 0x4}),
    ('no stores (0x8)', {'chain_dbg_skip': 0x8}),
    ('sa=5', {'chain_sa': 5}),
    ('sa=7', {'chain_sa': 7}),
    ('end', {}),
]
for name, knobs in SETS[:-1]:
    for k_, v_ in (('chain_dbg_skip', 0), ('chain_nsub', 2), ('chain_sa', 0)):
        tc.tuning_set(k_, v_)
    for k_, v_ in knobs.items():
        tc.tuning_set(k_, v_)
    t = timed(lambda: tc.conv_chain_single(x, cin, None, 0, nbr, n, ks ** 3, wt, cout))
    print(f'{name:50s} {t:8.1f} us')
