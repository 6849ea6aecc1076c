"""Which steps are slow, and is the time lost on the host or on the device?  (diagnostic)"""
import gc
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openscene_b200 import engine, matching, synth  # noqa: E402

dev = torch.device('cuda:0')
coords = torch.from_numpy(synth.scene('config2_200k')).to(dev)
feats = torch.ones(coords.shape[0], 3, device=dev)
text = torch.from_numpy(synth.text_embeddings(20)).to(dev)
eng = engine.FusedMinkUNet(synth.build_model('MinkUNet34C', 768, seed=0).eval().to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
step = lambda: matching._scores(eng(coords, feats), None, text, normalize=True)
for _ in range(10):
    step()
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else 'plain'
gc.collect(); gc.disable()
K = 200
ev, host = [], []
stats0 = torch.cuda.memory_stats()
t_prev = time.perf_counter()
for i in range(K):
    if mode != 'noflush':
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); step(); b.record()
    t1 = time.perf_counter()
    ev.append((a, b)); host.append((t0 - t_prev, t1 - t0)); t_prev = t1
torch.cuda.synchronize()
stats1 = torch.cuda.memory_stats()
ms = [a.elapsed_time(b) for a, b in ev]
srt = sorted(ms)
print(f'mode={mode} mean {sum(ms) / K:.3f} median {srt[K // 2]:.3f} max {srt[-1]:.2f}  cudaMalloc calls during loop: '
      f"{stats1['num_device_alloc'] - stats0['num_device_alloc']}  frees: {stats1['num_device_free'] - stats0['num_device_free']}")
for i, m in enumerate(ms):
    if m > 2 * srt[K // 2]:
        print(f'  step {i:3d}: gpu {m:7.2f} ms | host in-step {1e3 * host[i][1]:7.2f} ms, host gap before {1e3 * host[i][0]:6.2f} ms')
