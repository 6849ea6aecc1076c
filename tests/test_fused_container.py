"""The fixed-stride fused-feature container (openscene_b200/fused_container.py): what arrives on the
device after `read_remapped` must be exactly what the reference's own FusedFeatureLoader.__getitem__ returned for the same
scene (tests/golden/loader_*.npz), and exactly what the oracle restatement of the remap computes at scale."""
import numpy as np
import pytest
import torch

from openscene_b200 import fused_container as fc
from oracle import loader_ref, voxelize_ref
from tests.util import golden


@pytest.mark.parametrize('case', ['train', 'val', 'train_legacy'])
def test_container_reproduces_reference_loader(case, tmp_path):
    g = golden(f'loader_{case}.npz')
    src, dst = str(tmp_path / 'scene.pt'), str(tmp_path / 'scene.osbf')
    d = {'feat': torch.from_numpy(g['feat']), 'mask_full': torch.from_numpy(g['mask_full'])}
    if 'legacy_mask' in g.files:
        d['mask'] = torch.from_numpy(g['legacy_mask'])
    torch.save(d, src)                                            # the reference's on-disk form (fusion_util.py:86-88)
    fc.convert_torch_save(src, dst)
    _, inds, _, _ = voxelize_ref.voxelize(g['locs'].astype(np.float64), g['matrix'])
    f = fc.FusedFeatureFile(dst)
    assert f.n_points == 3000 and f.channels == 16
    feat, mask = f.read_remapped(inds, str(g['split']), device='cpu')
    assert mask.dtype == torch.bool and np.array_equal(mask.numpy(), g['mask'])
    assert feat.dtype == torch.float16 and np.array_equal(feat.numpy(), g['feat_3d'])


@pytest.mark.parametrize('c', [768, 20])                          # 20 channels: 40-byte rows padded to a 48-byte stride
def test_container_round_trip_and_remap_at_scale(c, tmp_path):
    g = torch.Generator().manual_seed(3)
    n_pts, n_vox = 200_000, 80_000
    mask_full = torch.rand(n_pts, generator=g) < 0.3
    feat = torch.randn(int(mask_full.sum()), c, generator=g).half()
    path = fc.write_container(str(tmp_path / 'a.osbf'), feat, mask_full)
    f = fc.FusedFeatureFile(path)
    assert (f.n_points, f.n_rows, f.channels) == (n_pts, feat.shape[0], c) and f.row_stride % 16 == 0
    all_feat, all_mask = f.read_all('cpu')
    assert torch.equal(all_feat, feat) and torch.equal(all_mask, mask_full)
    vox_ind = torch.randperm(n_pts, generator=g)[:n_vox]          # representative points are not sorted (voxel order)
    for split in ('train', 'val'):
        got_f, got_m = f.read_remapped(vox_ind, split, device='cpu')
        ref_f, ref_m = loader_ref.remap_fused_features(feat, mask_full, vox_ind, split)
        assert torch.equal(got_m, ref_m) and torch.equal(got_f, ref_f)
    rows, mk = f.rows_for(vox_ind.numpy())
    assert rows.size == int(mk.sum()) < n_vox                     # only the kept rows are read from the file
    e_f, e_m = f.read_remapped(torch.zeros(0, dtype=torch.int64), 'train', device='cpu')
    assert e_f.shape == (0, c) and e_m.numel() == 0


def test_container_rejects_bad_input(tmp_path):
    feat = torch.zeros(5, 16).half()
    with pytest.raises(ValueError, match='set entries'):
        fc.write_container(str(tmp_path / 'b.osbf'), feat, torch.ones(7, dtype=torch.bool))
    p = str(tmp_path / 'c.osbf')
    open(p, 'wb').write(b'not a container' * 10)
    with pytest.raises(ValueError, match='bad magic'):
        fc.FusedFeatureFile(p)
    good = fc.write_container(str(tmp_path / 'd.osbf'), feat, torch.tensor([1, 0, 1, 1, 0, 1, 1], dtype=torch.bool))
    raw = open(good, 'rb').read()
    open(p, 'wb').write(raw[:-40])                                # truncated rows
    with pytest.raises(ValueError, match='truncated'):
        fc.FusedFeatureFile(p)
    f = fc.FusedFeatureFile(good)
    with pytest.raises(ValueError, match='out of range'):
        f.rows_for(np.array([0, 7]))
    with pytest.raises(ValueError, match='out of range'):
        f.read_rows(np.array([5]), device='cpu')


@pytest.mark.gpu
def test_container_to_device_equals_remap_kernel(tmp_path):
    """File -> device through the container == the two-step path (whole dict resident, csrc/remap.cu) on the GPU."""
    from openscene_b200.fused_features import remap_fused_features
    g = torch.Generator().manual_seed(5)
    n_pts, n_vox, c = 500_000, 200_000, 768
    mask_full = torch.rand(n_pts, generator=g) < 0.5
    feat = torch.randn(int(mask_full.sum()), c, generator=g).half()
    f = fc.FusedFeatureFile(fc.write_container(str(tmp_path / 'g.osbf'), feat, mask_full))
    vox_ind = torch.randperm(n_pts, generator=g)[:n_vox]
    for split in ('train', 'val'):
        a_f, a_m = f.read_remapped(vox_ind.cuda(), split, device='cuda:0')
        b_f, b_m = remap_fused_features(feat, mask_full, vox_ind, split, device='cuda:0')
        assert a_f.is_cuda and torch.equal(a_f, b_f) and torch.equal(a_m, b_m)
