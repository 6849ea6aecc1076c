"""A handful of launches of one level-0 convolution through the persistent kernel (for `ncu --set full -k regex:k_conv_chain`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openscene_b200 import synth, tc  # noqa: E402
from openscene_b200.coords import CoordinateManager  # noqa: E402

dev = torch.device('cuda:0')
cin, cout, ks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
coords = torch.from_numpy(synth.scene('config2_200k')).to(dev)
cm = CoordinateManager(coords)
n = cm.sets[1].n
nbr, K = (cm.kernel_map(1, 1, ks).nbr, ks ** 3) if ks > 1 else (None, 1)
x = tc.to_split(torch.randn(n, cin, device=dev))
wt = tc.pack_weight_tiles(torch.randn(K, cin, cout, device=dev) * 0.05)
sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    tc.conv_chain_single(x, cin, None, 0, nbr, n, K, wt, cout, sc, sh, None, True, True, False, None)
torch.cuda.synchronize()
print('done')
