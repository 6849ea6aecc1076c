"""CPU restatement of the fused-feature remap inside ``FusedFeatureLoader.__getitem__``.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ``dataset/feature_loader.py:101-172``.  PINNED: ``tests/golden/loader_*.npz`` hold what the reference's own
``FusedFeatureLoader`` returned for synthetic ``.pth`` / ``.pt`` files written to a scratch directory
(``scripts/make_golden.py``; ``SharedArray`` stubbed, ``torch.load`` defaulting to ``weights_only=False`` as the
reference's PyTorch did).
"""
import torch


def remap_fused_features(feat_3d, mask_full, vox_ind, split='train', legacy_mask=None):
    """feat_3d [M,C]; mask_full bool [N_pts]; vox_ind int64 [N_vox] (the voxeliser's representative points);
    legacy_mask: the 'mask' entry of the old three-key files (bool [M]).
    train: voxelisation runs over ALL points -> (feat rows of voxels that have a feature, mask per voxel).
    val/test with the two-key format: the caller voxelises all points too and every voxel keeps a row (zeros if none).
    Returns (feat_3d_vox, mask_vox)."""
    mask_chunk = torch.as_tensor(mask_full).clone()
    vox_ind = torch.as_tensor(vox_ind).long()
    if legacy_mask is not None:                                   # :112-115, 146-149
        feat_3d = feat_3d[legacy_mask]
        mask_chunk[mask_chunk.clone()] = legacy_mask
    if split != 'train':                                          # :107-111, 165-170
        assert legacy_mask is None
        full = torch.zeros((mask_chunk.shape[0], feat_3d.shape[1]), dtype=feat_3d.dtype)
        full[mask_chunk] = feat_3d
        return full[vox_ind], mask_chunk[vox_ind]
    mask = mask_chunk[vox_ind]                                    # :127
    mask_ind = mask_chunk.nonzero(as_tuple=False)[:, 0]
    index1 = -torch.ones(mask_chunk.shape[0], dtype=int)
    index1[mask_ind] = mask_ind
    index1 = index1[vox_ind]
    chunk_ind = index1[index1 != -1]
    index2 = torch.zeros(mask_chunk.shape[0])
    index2[mask_ind] = 1
    index3 = torch.cumsum(index2, dim=0, dtype=int)
    return feat_3d[index3[chunk_ind] - 1], mask                   # :140-145
