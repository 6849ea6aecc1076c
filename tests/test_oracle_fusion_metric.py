"""Restatements of the 8f "next" rows (multi-view fusion mapping, segmentation metrics) against vectors produced by the
REFERENCE's own code (scripts/make_golden.py: fusion_util.PointCloudToImageMapper, util/metric.py, util/util.py).
Bit-exact: integer / index work."""
import numpy as np
import pytest
import torch

from oracle import fusion_ref, metric_ref
from openscene_b200.synth import fusion_case
from tests.util import golden

FUSION_CASES = ['depth_cut10', 'depth_cut0', 'nodepth_cut5']
METRIC_CASES = ['scannet20', 'mp160_nofeat', 'nuscenes16']


@pytest.mark.parametrize('case', FUSION_CASES)
def test_mapping_restatement_matches_reference(case):
    g = golden(f'fusion_mapping_{case}.npz')
    pts, poses, depths, intr = fusion_case(int(g['seed']), int(g['n']), bool(g['with_depth']))
    for f, (pose, depth) in enumerate(zip(poses, depths)):
        m = fusion_ref.compute_mapping(pose, pts, depth, intr, (320, 240), int(g['cut']))
        assert np.array_equal(m, g['mapping'][f])
        assert m[:, 2].sum() > 100          # the case exercises the visible branch ...
        assert (m[:, 2] == 0).sum() > 100   # ... and the rejected one


def test_fuse_frames_is_a_running_mean():
    pts, poses, depths, intr = fusion_case(5, 3000, True)
    rng = np.random.RandomState(0)
    feats = [torch.from_numpy(rng.randn(240, 320, 16).astype(np.float16)) for _ in poses]
    bank, counter, ids = fusion_ref.fuse_frames(pts, poses, depths, feats, intr, (320, 240), 10)
    maps = [fusion_ref.compute_mapping(p, pts, d, intr, (320, 240), 10) for p, d in zip(poses, depths)]
    vis = np.stack([m[:, 2] for m in maps], 1)
    assert np.array_equal(counter[:, 0].numpy(), vis.sum(1).astype(np.float32))
    assert np.array_equal(ids.numpy(), np.nonzero(vis.sum(1))[0])
    i = int(np.argmax(vis.sum(1)))
    want = sum(feats[f][maps[f][i, 0], maps[f][i, 1]].float() for f in range(len(poses)) if vis[i, f]) / float(vis[i].sum())
    assert torch.allclose(bank[i], want, rtol=1e-6, atol=1e-7)
    assert torch.all(bank[vis.sum(1) == 0] == 0)


@pytest.mark.parametrize('case', METRIC_CASES)
def test_metric_restatement_matches_reference(case):
    g = golden(f'metric_{case}.npz')
    C = int(g['C'])
    conf = metric_ref.confusion_matrix(g['pred'], g['gt'], C)
    assert np.array_equal(conf.astype(np.int64), g['confusion'])
    miou, _ = metric_ref.mean_iou(g['pred'], g['gt'], C)
    assert miou == pytest.approx(float(g['miou']), rel=1e-12)
    if 'inter' in g.files:
        i, u, t = metric_ref.intersection_and_union(g['pred'], g['gt'], C)
        assert np.array_equal(i, g['inter']) and np.array_equal(u, g['union']) and np.array_equal(t, g['target'])


@pytest.mark.parametrize('case', ['train', 'val', 'train_legacy'])
def test_loader_remap_restatement_matches_reference(case):
    """oracle voxeliser + oracle remap == what the reference's FusedFeatureLoader.__getitem__ returned."""
    from oracle import loader_ref, voxelize_ref
    g = golden(f'loader_{case}.npz')
    cv, inds, inv, _ = voxelize_ref.voxelize(g['locs'], g['matrix'])
    assert np.array_equal(cv.astype(np.int32), g['coords'][:, 1:]) and np.all(g['coords'][:, 0] == 1)
    legacy = torch.from_numpy(g['legacy_mask']) if 'legacy_mask' in g.files else None
    feat, mask = loader_ref.remap_fused_features(torch.from_numpy(g['feat']), torch.from_numpy(g['mask_full']), inds, str(g['split']), legacy)
    assert np.array_equal(mask.numpy(), g['mask'])
    assert feat.dtype == torch.float16 and np.array_equal(feat.numpy(), g['feat_3d'])
    if 'inds_reverse' in g.files:
        assert np.array_equal(inv, g['inds_reverse'])


def test_intrinsics_helpers_match_reference():
    """make_intrinsic / adjust_intrinsic (host-side NumPy in openscene_b200/fusion.py) against fusion_util.py:17-39."""
    from openscene_b200.fusion import adjust_intrinsic, make_intrinsic
    g = golden('fusion_intrinsics.npz')
    k0 = make_intrinsic(fx=577.870605, fy=577.870605, mx=319.5, my=239.5)
    assert np.array_equal(k0, g['k0'])
    assert np.array_equal(adjust_intrinsic(k0.copy(), [640, 480], (320, 240)), g['k1'])
    assert np.array_equal(adjust_intrinsic(make_intrinsic(1075.1, 1075.8, 629.7, 522.3), [1280, 1024], (640, 512)), g['k2'])
    same = make_intrinsic(1.0, 2.0, 3.0, 4.0)
    assert adjust_intrinsic(same, [320, 240], [320, 240]) is same


def test_dataset_class_counts_match_reference_label_lists():
    """evaluate() picks the class count from the dataset name (util/metric.py:47-60, dataset/label_constants.py)."""
    from openscene_b200.metric import _DATASET_CLASSES
    g = golden('metric_class_counts.npz')
    for name, count in zip(g['names'].tolist(), g['counts'].tolist()):
        picked = next(n for key, n in _DATASET_CLASSES if key in name)       # first match wins, as the reference's elif chain
        assert picked == count, (name, picked, count)
