"""Repeated launches of one 3^3 convolution on a few CTAs: which rows are wrong / differ between runs (plain and scattered output)?
    python scripts/chain_debug3.py grid cin cout ks [runs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import synth, tc  # noqa: E402
from openscene_b200.coords import CoordinateManager  # noqa: E402

grid, cin, cout, ks = (int(v) for v in sys.argv[1:5])
runs = int(sys.argv[5]) if len(sys.argv) > 5 else 8
knobs = dict(kv.split('=') for kv in sys.argv[6:])
dev = torch.device('cuda:0')
c = torch.from_numpy(synth.scene('tiny')).to(dev)
cm = CoordinateManager(c)
n = cm.sets[1].n
nbr = cm.kernel_map(1, 1, ks).nbr
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, cin, device=dev, generator=g)
w = torch.randn(ks ** 3, cin, cout, device=dev, generator=g) * 0.05
xs, wt = tc.to_split(x), tc.pack_weight_tiles(w)
tc.tuning_set('chain_grid', grid)
for k_, v_ in knobs.items():
    tc.tuning_set(k_, int(v_, 0))
ref = torch.zeros(n, cout, device=dev, dtype=torch.float64)
for k in range(ks ** 3):
    o = (nbr[k] >= 0).nonzero()[:, 0]
    ref[o] += x.double()[nbr[k][o].long()] @ w[k].double()
perm = torch.randperm(n, device=dev).int()
m_tiles = (n + 127) // 128
print(f'n={n} m_tiles={m_tiles} grid={grid}: units per CTA', [m_tiles * (b + 1) // grid - m_tiles * b // grid for b in range(grid)])
outs = []
for rep in range(runs):
    scat = rep % 3 == 2
    _, o = tc.conv_chain_single(xs, cin, None, 0, nbr, n, ks ** 3, wt, cout, None, None, None, False, False, True, perm if scat else None)
    torch.cuda.synchronize()
    o = o[perm.long()] if scat else o
    outs.append(o.clone())
    err = ((o.double() - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-9))
    bad = (err > 1e-4).nonzero()[:, 0]
    diff0 = (o != outs[0]).any(dim=1).nonzero()[:, 0]
    print(f'run: max rel err {float(err.max()):.3e}, rows off vs fp64 {len(bad)} tiles {sorted(set((bad // 128).tolist()))[:12]}; '
          f'rows differing from run 0: {len(diff0)} tiles {sorted(set((diff0 // 128).tolist()))[:12]}')
    if len(diff0):
        r = int(diff0[0])
        d = (o[r] - outs[0][r])
        print('   first differing row', r, '(row in tile', r % 128, ') cols differing', int((d != 0).sum()), 'max abs diff', float(d.abs().max()), 'row norm', float(outs[0][r].norm()))
