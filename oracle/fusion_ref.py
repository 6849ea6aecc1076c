"""CPU restatement of the multi-view feature-fusion accumulate.
TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``compute_mapping`` follows ``PointCloudToImageMapper.compute_mapping``
  (``scripts/feature_fusion/fusion_util.py:102-139``) and is PINNED: ``tests/golden/fusion_mapping_*.npz`` hold the
  outputs of the reference's own class (imported with a stub ``tensorflow`` module, which that method never touches;
  ``scripts/make_golden.py``).
* ``fuse_frames`` follows the accumulation loop of ``process_one_scene``
  (``scripts/feature_fusion/scannet_openseg.py:74-108``): per frame, in frame order, ``counter[mask] += 1`` and
  ``sum_features[mask] += feat_2d[:, v, u]`` (fp32 accumulator, fp16 pixel feature), then
  ``counter[counter == 0] = 1e-5; feat_bank = sum_features / counter`` and ``point_ids`` = points seen by >= 1 frame.
  That function reads images and poses from disk and calls OpenSeg, so it cannot be run here; the restatement takes
  the per-frame tensors instead.
"""
import numpy as np
import torch


def compute_mapping(camera_to_world, coords, depth, intrinsic, image_dim, cut_bound=0, vis_thres=0.25):
    """fusion_util.py:102-139.  coords [N,3]; depth [H,W] float (metres) or None; intrinsic >= 3x3;
    image_dim = (W, H).  Returns int [N,3] = (row v, col u, visible)."""
    n = coords.shape[0]
    mapping = np.zeros((3, n), dtype=int)
    homog = np.concatenate([coords, np.ones([n, 1])], axis=1).T            # [4,N] float64
    p = np.matmul(np.linalg.inv(camera_to_world), homog)
    with np.errstate(all='ignore'):
        p[0] = (p[0] * intrinsic[0][0]) / p[2] + intrinsic[0][2]
        p[1] = (p[1] * intrinsic[1][1]) / p[2] + intrinsic[1][2]
        pi = np.round(p).astype(int)                                       # round half to even, then truncate to int64
    inside = (pi[0] >= cut_bound) * (pi[1] >= cut_bound) * (pi[0] < image_dim[0] - cut_bound) * (pi[1] < image_dim[1] - cut_bound)
    if depth is not None:
        d = depth[pi[1][inside], pi[0][inside]]
        inside[inside == True] = np.abs(d - p[2][inside]) <= vis_thres * d   # noqa: E712 (reference idiom)
    else:
        inside = (p[2] > 0) * inside
    mapping[0][inside] = pi[1][inside]
    mapping[1][inside] = pi[0][inside]
    mapping[2][inside] = 1
    return mapping.T


def fuse_frames(points, poses, depths, feats_hwc, intrinsic, image_dim, cut_bound, vis_thres=0.25):
    """scannet_openseg.py:74-108.  poses: list of 4x4 camera-to-world; depths: list of [H,W] float64 (or None);
    feats_hwc: list of fp16 tensors [H,W,C] (the reference holds the same memory as a [C,H,W] permuted view).
    Returns (feat_bank fp32 [N,C], counter fp32 [N,1] before the 1e-5 patch, point_ids int64)."""
    n, c = points.shape[0], feats_hwc[0].shape[-1]
    counter = torch.zeros((n, 1))
    sum_features = torch.zeros((n, c))
    seen = torch.zeros(n, dtype=torch.bool)
    for pose, depth, feat in zip(poses, depths, feats_hwc):
        m = torch.from_numpy(compute_mapping(pose, points, depth, intrinsic, image_dim, cut_bound, vis_thres))
        mask = m[:, 2] != 0
        if mask.sum() == 0:
            continue
        seen |= mask
        feat_2d_3d = feat[m[:, 0], m[:, 1], :]                             # == feat_2d[:, v, u].permute(1, 0)
        counter[mask] += 1
        sum_features[mask] += feat_2d_3d[mask]
    raw_counter = counter.clone()
    counter[counter == 0] = 1e-5
    return sum_features / counter, raw_counter, torch.nonzero(seen)[:, 0]
