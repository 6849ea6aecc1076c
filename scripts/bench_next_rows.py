"""Measure the SURVEY 8f 'next' rows on the GPU: multi-view fusion accumulate and the segmentation metrics.
Prints one JSON line per row (CUDA-event timed, warm, inputs resident in HBM) with the HBM roofline fraction and a
CPU-oracle baseline on a bounded sample.  Not part of bench.py's contract (that is the north-star path)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openscene_b200 import _cabi as C                                 # noqa: E402
from openscene_b200.fusion import FeatureFusion, PointCloudToImageMapper   # noqa: E402
from openscene_b200 import metric                                     # noqa: E402
from openscene_b200.synth import fusion_case                          # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def timed(fn, reps=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def bench_fusion(n=1_000_000, c=768, F=32):
    from oracle import fusion_ref
    pts, poses, depths, intr = fusion_case(77, n, True, n_frames=F)
    mapper = PointCloudToImageMapper(image_dim=(320, 240), intrinsics=intr, cut_bound=10, device='cuda')
    feats = (torch.randn(F, 240, 320, c, device='cuda') * 0.5).half()
    fuser = FeatureFusion(pts, c, mapper)
    d_dev = torch.stack([torch.from_numpy(d) for d in depths]).cuda()
    l0 = C.lib().osb_launch_count()
    fuser.add_frames(poses, list(d_dev), feats)
    launches = C.lib().osb_launch_count() - l0
    ms = timed(lambda: fuser.add_frames(poses, list(d_dev), feats))
    maps = torch.stack([mapper.compute_mapping(p, pts, d, as_tensor=True) for p, d in zip(poses, depths)])
    vis = maps[:, :, 2].long()
    pairs, touched = int(vis.sum()), int((vis.sum(0) > 0).sum())
    # algorithmic bytes: points once per frame (24 B), depth probe (8 B) per in-image pair ~ per pair, the pixel feature
    # (2C) per visible pair, the fp32 sum row read + written once per touched point (8C), counter (8 B)
    alg = 24 * n + 8 * pairs + 2 * c * pairs + 8 * c * touched + 8 * touched
    peak, src = peak_gbs()
    # CPU oracle on a bounded sample: 4 frames, 100k points
    ns, Fs = 100_000, 4
    t0 = time.time()
    fusion_ref.fuse_frames(pts[:ns], poses[:Fs], depths[:Fs], [f.cpu() for f in feats[:Fs]], intr, (320, 240), 10)
    cpu_s = time.time() - t0
    print(json.dumps({'row': '8f-2 fusion accumulate', 'points': n, 'frames_per_call': F, 'feat_dim': c, 'ms_per_call': ms,
                      'point_frames_per_s': n * F / ms * 1e3, 'visible_pairs': pairs, 'touched_points': touched,
                      'gpu_launches_per_call': int(launches),
                      'roofline': {'bound': 'hbm', 'achieved': alg / ms / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': alg / ms / 1e6 / peak,
                                   'peak_source': src, 'algorithmic_bytes': alg},
                      'cpu_baseline': {'value': ns * Fs / cpu_s, 'unit': 'point-frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                       'sample': f'{ns} points x {Fs} frames, {cpu_s:.1f} s'}}))


def bench_metric(n=20_000_000, Cn=20):
    from oracle import metric_ref
    g = torch.Generator(device='cuda').manual_seed(0)
    gt = torch.randint(0, Cn, (n,), device='cuda', generator=g)
    gt[torch.rand(n, device='cuda', generator=g) < 0.1] = 255
    pred = torch.where(torch.rand(n, device='cuda', generator=g) < 0.6, gt.clamp(max=Cn - 1), torch.randint(0, Cn, (n,), device='cuda', generator=g))
    meter = metric.ConfusionMeter(Cn)
    ms = timed(lambda: meter.update(pred, gt))
    ms_iu = timed(lambda: metric.intersectionAndUnionGPU(pred, gt, Cn, 255))
    peak, src = peak_gbs()
    alg = 16 * n
    ns = 2_000_000
    p, q = pred[:ns].cpu().numpy(), gt[:ns].cpu().numpy()
    t0 = time.time(); metric_ref.confusion_matrix(p, q, Cn); cpu_s = time.time() - t0
    print(json.dumps({'row': '8f-4 confusion matrix / intersection-union', 'labels': n, 'classes': Cn, 'dtype': 'int64',
                      'confusion_ms': ms, 'inter_union_ms': ms_iu, 'labels_per_s': n / ms * 1e3,
                      'roofline': {'bound': 'hbm', 'achieved': alg / ms / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': alg / ms / 1e6 / peak,
                                   'peak_source': src, 'algorithmic_bytes': alg},
                      'cpu_baseline': {'value': ns / cpu_s, 'unit': 'labels/s', 'cores': 1, 'kind': 'port', 'sample': f'{ns} labels, {cpu_s:.2f} s'}}))


if __name__ == '__main__':
    bench_fusion()
    bench_metric()
