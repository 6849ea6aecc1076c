"""Host-side enqueue time of one engine step (no device sync inside the timed host region except the ones the
coordinate manager needs), split by phase.  Tuning aid."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openscene_b200 import engine, matching, synth  # noqa: E402
from openscene_b200.coords import CoordinateManager  # noqa: E402

dev = torch.device('cuda:0')
coords = torch.from_numpy(synth.scene('config2_200k')).to(dev)
feats = torch.ones(coords.shape[0], 3, device=dev)
text = torch.from_numpy(synth.text_embeddings(20)).to(dev)
model = synth.build_model('MinkUNet34C', 768, seed=0).eval().to(dev)
eng = engine.FusedMinkUNet(model)
for _ in range(3):
    matching._scores(eng(coords, feats), None, text, normalize=True)
torch.cuda.synchronize()
R = 20
PYR = int(os.environ.get('PYR', '4'))
t_cm = t_fw = t_mt = t_b = 0.0
t0 = time.perf_counter()
for _ in range(R):
    a = time.perf_counter()
    cm = CoordinateManager(coords, pyramid_levels=PYR)
    torch.cuda.synchronize(); a1 = time.perf_counter(); t_b += a1 - a
    ts = 1
    for _ in range(4):
        ts = cm.stride(ts, 2)
    for t in (1, 2, 4, 8, 16):
        cm.kernel_map(t, t, 3)
    for t in (1, 2, 4, 8):
        cm.kernel_map(t, 2 * t, 2).transposed()
    b = time.perf_counter()
    out = eng(coords, feats, coordinate_manager=cm)
    c = time.perf_counter()
    matching._scores(out, None, text, normalize=True)
    d = time.perf_counter()
    t_cm += b - a; t_fw += c - b; t_mt += d - c
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f'PYR={PYR} build+sync {1e3 * t_b / R:.2f} ms;', end=' ')
print(f'per step: wall {1e3 * tot / R:.2f} ms | host: coordinate phase (incl. its syncs) {1e3 * t_cm / R:.2f} ms, '
      f'conv-chain enqueue {1e3 * t_fw / R:.2f} ms, match enqueue {1e3 * t_mt / R:.3f} ms')
