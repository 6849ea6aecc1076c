"""Per-convolution device times of one engine forward (CUDA events around every native convolution call).
    python scripts/layer_times.py [workload] [arch]        env: OSB_CHAIN, OSB_CHAIN_MAX_TILES, knobs via KNOBS="name=value,..."
With OSB_CHAIN=1 use OSB_CHAIN_MAX_TILES=0 so that every layer is its own launch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import _cabi, engine, synth, tc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'config2_200k'
arch = sys.argv[2] if len(sys.argv) > 2 else 'MinkUNet34C'
for kv in filter(None, os.environ.get('KNOBS', '').split(',')):
    k, v = kv.split('=')
    tc.tuning_set(k, int(v))
dev = torch.device('cuda:0')
coords = torch.from_numpy(synth.scene(workload)).to(dev)
feats = torch.ones(coords.shape[0], 3, device=dev)
model = synth.build_model(arch, 768, seed=0).eval().to(dev)
eng = engine.FusedMinkUNet(model)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    eng(coords, feats)
names = ('osb_conv_fwd_tc', 'osb_convtr_fwd_tc', 'osb_conv_chain_launch')
pend = []
orig_lib = _cabi.lib


def make_hook(fn):
    def hooked(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a); e1.record()
        pend.append((e0, e1))
        return r
    return hooked


hooks = {nm: make_hook(getattr(orig_lib(), nm)) for nm in names}


class Proxy:
    def __getattr__(self, name):
        return hooks[name] if name in hooks else getattr(orig_lib(), name)


reps = 5
acc = None
for r in range(reps):
    pend.clear()
    eng.layer_log = []
    flush.zero_()
    _cabi.lib = lambda: Proxy()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng(coords, feats); b.record()
    _cabi.lib = orig_lib
    torch.cuda.synchronize()
    ts = [x.elapsed_time(y) * 1e3 for x, y in pend]
    acc = ts if acc is None else [min(p, q) for p, q in zip(acc, ts)]
    total = a.elapsed_time(b)
log = eng.layer_log
print(f'# {workload} {arch} OSB_CHAIN={os.environ.get("OSB_CHAIN", "1")} MAX_TILES={os.environ.get("OSB_CHAIN_MAX_TILES", "-1")} '
      f'KNOBS={os.environ.get("KNOBS", "")}: forward {total:.3f} ms, conv launches {len(acc)}, sum {sum(acc) / 1e3:.3f} ms (min of {reps} reps, us)')
if len(log) == len(acc):
    for (n, K, cin, cout, tag), t in zip(log, acc):
        print(f'{n:8d} K={K:2d} {cin:4d}->{cout:4d} {tag:9s} {t:8.1f}')
else:
    print('groups:', ' '.join(f'{t:.1f}' for t in acc), f'({len(log)} layers)')
