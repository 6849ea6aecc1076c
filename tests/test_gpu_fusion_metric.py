"""The 8f "next" rows on the GPU: multi-view feature fusion (scripts/feature_fusion) and the segmentation metrics
(util/metric.py, util/util.py) against the reference's own outputs (tests/golden, scripts/make_golden.py) and the
oracle restatements.  Everything here is bit-exact: index work, integer counting, and fp32 sums taken in the
reference's order."""
import numpy as np
import pytest
import torch

from openscene_b200.synth import fusion_case
from tests.util import golden

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
IMG = (320, 240)


@pytest.mark.parametrize('case', ['depth_cut10', 'depth_cut0', 'nodepth_cut5'])
def test_mapping_matches_reference_vectors(case):
    from openscene_b200.fusion import PointCloudToImageMapper
    g = golden(f'fusion_mapping_{case}.npz')
    pts, poses, depths, intr = fusion_case(int(g['seed']), int(g['n']), bool(g['with_depth']))
    mapper = PointCloudToImageMapper(image_dim=IMG, intrinsics=intr, visibility_threshold=0.25, cut_bound=int(g['cut']), device=DEV)
    for f, (pose, depth) in enumerate(zip(poses, depths)):
        m = mapper.compute_mapping(pose, pts, depth, as_tensor=True)
        assert m.dtype == torch.int32 and m.shape == (len(pts), 3)
        assert np.array_equal(m.cpu().numpy(), g['mapping'][f])
        mn = mapper.compute_mapping(pose, pts, depth)                    # default: NumPy int array like the reference
        assert isinstance(mn, np.ndarray) and mn.dtype == np.zeros(1, dtype=int).dtype and np.array_equal(mn, g['mapping'][f])
        buf = np.zeros((len(pts), 4), dtype=int)
        buf[:, 1:4] = mn                                                 # the reference's own use (scannet_openseg.py:95)
    # float32 points are promoted to float64 exactly as np.concatenate([coords, ones]) does
    m32 = mapper.compute_mapping(poses[0], pts.astype(np.float32), depths[0])
    from oracle import fusion_ref
    want = fusion_ref.compute_mapping(poses[0], pts.astype(np.float32), depths[0], intr, IMG, int(g['cut']))
    assert np.array_equal(m32, want)


@pytest.mark.parametrize('n,c,n_frames,image_dim,with_depth,seed', [
    (6000, 768, 3, (320, 240), True, 41),        # the reference's shape: OpenSeg 768-d, 320x240
    (5000, 512, 4, (320, 240), False, 42),       # LSeg width, no depth image (nuScenes branch)
    (4000, 64, 40, (160, 120), True, 43),        # more than one native batch of 32 frames
    (37, 8, 2, (160, 120), True, 44),            # ragged tiny case
])
def test_fuse_frames_bit_exact(n, c, n_frames, image_dim, with_depth, seed):
    from openscene_b200.fusion import FeatureFusion, PointCloudToImageMapper
    from oracle import fusion_ref
    pts, poses, depths, intr = fusion_case(seed, n, with_depth, image_dim=image_dim, n_frames=n_frames)
    W, H = image_dim
    g = torch.Generator().manual_seed(seed)
    feats = [(torch.randn(H, W, c, generator=g) * 0.5).half() for _ in range(n_frames)]
    bank_ref, counter_ref, ids_ref = fusion_ref.fuse_frames(pts, poses, depths, feats, intr, image_dim, 10)
    mapper = PointCloudToImageMapper(image_dim=image_dim, intrinsics=intr, visibility_threshold=0.25, cut_bound=10, device=DEV)
    fuser = FeatureFusion(pts, c, mapper)
    half = n_frames // 2                          # two calls: the running state carries over
    fuser.add_frames(poses[:half], depths[:half] if with_depth else None, feats[:half])
    # second half through the reference's [C,H,W] permuted view
    fuser.add_frames(poses[half:], depths[half:] if with_depth else None, torch.stack(feats[half:]).permute(0, 3, 1, 2))
    bank, ids = fuser.finalize()
    assert counter_ref.sum() > 0
    assert torch.equal(fuser.counter.cpu(), counter_ref[:, 0])
    assert torch.equal(ids.cpu(), ids_ref)
    assert torch.equal(bank.cpu(), bank_ref)      # same fp32 additions in the same order, IEEE division


def test_fusion_rejects_bad_arguments():
    from openscene_b200.fusion import FeatureFusion, PointCloudToImageMapper
    pts, poses, depths, intr = fusion_case(1, 100, True, image_dim=(160, 120), n_frames=1)
    mapper = PointCloudToImageMapper(image_dim=(160, 120), intrinsics=intr, cut_bound=0, device=DEV)
    fuser = FeatureFusion(pts, 12, mapper)        # width not a multiple of 8
    with pytest.raises(RuntimeError, match='multiple of 8'):
        fuser.add_frames(poses, depths, torch.zeros(1, 120, 160, 12, dtype=torch.float16))


@pytest.mark.parametrize('case', ['scannet20', 'mp160_nofeat', 'nuscenes16'])
@pytest.mark.parametrize('dtype', [torch.int32, torch.int64])
def test_metrics_match_reference_vectors(case, dtype):
    from openscene_b200 import metric
    g = golden(f'metric_{case}.npz')
    C = int(g['C'])
    pred, gt = torch.from_numpy(g['pred']).to(dtype), torch.from_numpy(g['gt']).to(dtype)
    conf = metric.confusion_matrix(pred, gt, C)
    assert conf.dtype == np.ulonglong and np.array_equal(conf.astype(np.int64), g['confusion'])
    ds = {'scannet20': 'scannet_3d', 'mp160_nofeat': 'matterport_3d_160', 'nuscenes16': 'nuscenes_3d'}[case]
    assert metric.evaluate(pred.to(DEV), gt.to(DEV), dataset=ds) == pytest.approx(float(g['miou']), rel=1e-12)
    # batch-wise accumulation on the device == one shot
    meter = metric.ConfusionMeter(C, device=DEV)
    for a in range(0, len(pred), 7001):
        meter.update(pred[a:a + 7001].to(DEV), gt[a:a + 7001].to(DEV))
    assert np.array_equal(meter.confusion().astype(np.int64), g['confusion'])
    if 'inter' in g.files:
        o = pred.to(DEV)
        keep = o.clone()
        i, u, t = metric.intersectionAndUnionGPU(o, gt.to(DEV), C, 255)
        assert i.is_cuda and i.dtype == torch.float32
        assert np.array_equal(i.cpu().numpy(), g['inter'].astype(np.float32))
        assert np.array_equal(u.cpu().numpy(), g['union'].astype(np.float32))
        assert np.array_equal(t.cpu().numpy(), g['target'].astype(np.float32))
        assert torch.equal(o, keep)


def test_metric_edge_cases():
    from openscene_b200 import metric
    from oracle import metric_ref
    # every point ignored; empty input
    assert metric.confusion_matrix(torch.zeros(10, dtype=torch.int64), torch.full((10,), 255), 5).sum() == 0
    assert metric.confusion_matrix(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 5).sum() == 0
    with pytest.raises(ValueError):
        metric.confusion_matrix(torch.tensor([7]), torch.tensor([1]), 5)
    # many classes: the global-atomics path (bins do not fit shared memory)
    rng = np.random.RandomState(0)
    C = 300
    gt, pred = rng.randint(0, C, 100000), rng.randint(0, C, 100000)
    gt[::9] = 255
    pred[::13] = 256
    assert np.array_equal(metric.confusion_matrix(pred, gt, C), metric_ref.confusion_matrix(pred, gt, C))
    # predictions outside 0..K-1 are dropped from the histograms, as histc does
    o, t = torch.tensor([0, 1, 300, 2, 2], device=DEV), torch.tensor([0, 2, 1, 2, 255], device=DEV)
    i, u, a = metric.intersectionAndUnionGPU(o, t, 3, 255)
    ir, ur, ar = metric_ref.intersection_and_union(o.cpu().numpy(), t.cpu().numpy(), 3, 255)
    assert i.cpu().tolist() == ir.tolist() and u.cpu().tolist() == ur.tolist() and a.cpu().tolist() == ar.tolist()


def test_fusion_fullsize_properties():
    """BASELINE-size scene (1M points, 768-d, 8 frames): size-independent properties instead of a CPU run."""
    from openscene_b200.fusion import FeatureFusion, PointCloudToImageMapper
    n, c, F = 1_000_000, 768, 8
    pts, poses, depths, intr = fusion_case(77, n, True, n_frames=F)
    mapper = PointCloudToImageMapper(image_dim=IMG, intrinsics=intr, cut_bound=10, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    feats = (torch.randn(F, 240, 320, c, generator=g, device=DEV) * 0.5).half()
    fuser = FeatureFusion(pts, c, mapper)
    fuser.add_frames(poses, depths, feats)
    bank, ids = fuser.finalize()
    maps = torch.stack([mapper.compute_mapping(p, pts, d, as_tensor=True) for p, d in zip(poses, depths)]).long()     # [F,N,3]
    vis = maps[:, :, 2]
    assert torch.equal(fuser.counter, vis.sum(0).float())
    assert torch.equal(ids, torch.nonzero(vis.sum(0) > 0)[:, 0])
    assert int(vis.sum()) > 100_000
    once = torch.nonzero(vis.sum(0) == 1)[:, 0][:20000]               # seen by exactly one frame: mean == that pixel's feature
    f_of = vis[:, once].argmax(0)
    want = feats[f_of, maps[f_of, once, 0], maps[f_of, once, 1]].float()
    assert torch.equal(bank[once], want)
    assert torch.all(bank[vis.sum(0) == 0] == 0)
    # the running state is additive: a second pass over the same frames doubles counter, mean of identical terms
    twice = torch.nonzero(vis.sum(0) == 2)[:, 0][:20000]
    fa = vis[:, twice].float().argmax(0)
    fb = (F - 1) - vis[:, twice].flip(0).float().argmax(0)
    want2 = (feats[fa, maps[fa, twice, 0], maps[fa, twice, 1]].float() + feats[fb, maps[fb, twice, 0], maps[fb, twice, 1]].float()) / 2.0
    assert torch.equal(bank[twice], want2)


@pytest.mark.parametrize('case', ['train', 'val', 'train_legacy'])
def test_loader_remap_matches_reference_loader(case):
    """GPU voxeliser + GPU remap == what the reference's FusedFeatureLoader.__getitem__ returned (tests/golden)."""
    from openscene_b200.fused_features import remap_fused_features
    from openscene_b200.voxelize import voxelize_points
    g = golden(f'loader_{case}.npz')
    cv, inds, inv, _ = voxelize_points(torch.from_numpy(g['locs']).to(DEV), g['matrix'])
    assert np.array_equal(cv.cpu().numpy(), g['coords'][:, 1:])
    legacy = g['legacy_mask'] if 'legacy_mask' in g.files else None
    feat, mask = remap_fused_features(g['feat'], g['mask_full'], inds, str(g['split']), legacy, device=DEV)
    assert mask.dtype == torch.bool and np.array_equal(mask.cpu().numpy(), g['mask'])
    assert feat.dtype == torch.float16 and np.array_equal(feat.cpu().numpy(), g['feat_3d'])


def test_loader_remap_fullsize_and_errors():
    from openscene_b200.fused_features import remap_fused_features
    from oracle import loader_ref
    g = torch.Generator().manual_seed(9)
    n_pts, n_vox, c = 1_000_000, 600_000, 768
    mask_full = torch.rand(n_pts, generator=g) < 0.02                       # 20k supervised points, as the train chunks
    m = int(mask_full.sum())
    feat = torch.randn(m, c, generator=g).half()
    vox_ind = torch.randperm(n_pts, generator=g)[:n_vox]
    for split in ('train', 'val'):
        f, mk = remap_fused_features(feat, mask_full, vox_ind, split, device=DEV)
        fr, mr = loader_ref.remap_fused_features(feat, mask_full, vox_ind, split)
        assert torch.equal(mk.cpu(), mr) and torch.equal(f.cpu(), fr)
    with pytest.raises(RuntimeError, match='True entries'):
        remap_fused_features(feat[:-1], mask_full, vox_ind, 'train', device=DEV)
    with pytest.raises(RuntimeError, match='outside'):
        remap_fused_features(feat, mask_full, torch.tensor([0, n_pts]), 'train', device=DEV)
    f, mk = remap_fused_features(feat, mask_full, torch.zeros(0, dtype=torch.int64), 'train', device=DEV)
    assert f.shape == (0, c) and mk.numel() == 0
