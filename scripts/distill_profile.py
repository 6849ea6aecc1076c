"""Where the time of one distillation step goes (torch profiler, CUDA kernels by total time).  python scripts/distill_profile.py [arch]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import distill, synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else 'MinkUNet18A'
dev = torch.device('cuda:0')
coords_np = synth.scene('config2_200k')
n0 = len(coords_np)
g = torch.Generator().manual_seed(100)
mask = torch.zeros(n0, dtype=torch.bool)
mask[torch.randperm(n0, generator=g)[:20000]] = True
feat3d = (torch.randn(20000, 768, generator=g) * 0.3).half().to(dev)
coords, feats, mask = torch.from_numpy(coords_np).to(dev), torch.ones(n0, 3, device=dev), mask.to(dev)
torch.manual_seed(0)
model = synth.build_model(arch, 768, seed=0).train().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
step = lambda: distill.distill_step(model, opt, coords, feats, feat3d, mask, 'cosine', translate=True)
for _ in range(5):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record(); torch.cuda.synchronize()
print(f'# {arch}: {a.elapsed_time(b) / 5:.2f} ms per step (device-resident batch)')
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
tot = sum(r[2] for r in rows)
print(f'# CUDA kernel time in 3 steps: {tot / 1e3:.1f} ms')
for k, c, t in sorted(rows, key=lambda r: -r[2])[:32]:
    print(f'{t / 3e3:8.3f} ms/step {100 * t / tot:5.1f}%  x{c // 3:4d}  {k[:110]}')
