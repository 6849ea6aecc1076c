"""Per-role cycle accounting of the persistent convolution kernel for ONE shape (tuning knob chain_dbg_clock).
    python scripts/conv_prof.py [workload] cin cout ks [knob=value ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import _cabi as C, synth, tc  # noqa: E402
from openscene_b200.coords import CoordinateManager  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'config2_200k'
cin, cout, ks = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (96, 96, 3)))
knobs = dict(kv.split('=') for kv in sys.argv[5:])
dev = torch.device('cuda:0')
c = torch.from_numpy(synth.scene(workload)).to(dev)
cm = CoordinateManager(c)
n = cm.sets[1].n
nbr = cm.kernel_map(1, 1, ks).nbr if ks > 1 else None
g = torch.Generator(device=dev).manual_seed(0)
x = tc.to_split(torch.randn(n, cin, device=dev, generator=g))
w = torch.randn(ks ** 3, cin, cout, device=dev, generator=g) * 0.05
wt = tc.pack_weight_tiles(w)
for k_, v_ in knobs.items():
    tc.tuning_set(k_, int(v_, 0))
grid = C.lib().osb_conv_chain_grid()
run = lambda: tc.conv_chain_single(x, cin, None, 0, nbr, n, ks ** 3, wt, cout)
for _ in range(3):
    run()
buf = torch.zeros(grid * 32, dtype=torch.int64, device=dev)
tc.tuning_set('chain_dbg_clock', buf.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record()
torch.cuda.synchronize()
tc.tuning_set('chain_dbg_clock', 0)
d = buf.view(grid, 32).double().cpu()
print(f'# {workload} n={n} {cin}->{cout} k={ks} knobs={knobs}: {a.elapsed_time(b) * 1e3:.1f} us with accounting on')
tot = (d[:, 2] - d[:, 0])
print(f'CTA cycles: mean {tot.mean():.0f} max {tot.max():.0f}; set-up {(d[:, 1] - d[:, 0]).mean():.0f}')
rows = [('B producer', 4, ('wait emptyB', '-', '-')), ('issuer 0', 8, ('wait fullB', 'wait fullA', 'wait accEmpty')),
        ('issuer 1', 12, ('wait fullB', 'wait fullA', 'wait accEmpty')), ('gather producer (1 of 5 warps)', 16, ('wait emptyA', 'index loads land', '-')),
        ('epilogue (1 of 4 warps)', 20, ('wait accFull', '-', '-'))]
for name, base, labels in rows:
    w0, w1, w2, t = (d[:, base + i].mean().item() for i in range(4))
    parts = ', '.join(f'{lab} {v:.0f}' for lab, v in zip(labels, (w0, w1, w2)) if lab != '-')
    print(f'{name:22s} loop {t:9.0f}  | {parts} | own work {t - w0 - w1 - w2:9.0f}')
