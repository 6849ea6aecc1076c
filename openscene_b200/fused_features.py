"""Fused-feature remap on the GPU: what ``FusedFeatureLoader.__getitem__``
(``dataset/feature_loader.py:101-172``) does with ``{feat, mask_full}`` after voxelisation -- per-voxel feature mask and
the matching feature rows -- as one native call (csrc/remap.cu) instead of nonzero / cumsum / index passes on the CPU.
The on-disk container itself (``torch.save`` of that dict, ``fusion_util.py:87-89``) is unchanged."""
import ctypes

import torch

from . import _cabi as C


def remap_fused_features(feat_3d, mask_full, vox_ind, split='train', legacy_mask=None, device='cuda'):
    """feat_3d [M,C] (fp16 as stored; any dtype whose row size is a multiple of 16 bytes), mask_full bool [N_pts],
    vox_ind int64 [N_vox] = ``voxelize(..., return_ind=True)[-1]`` over ALL points.  Returns CUDA tensors
    (feat_3d_vox, mask_vox) equal to the loader's ``feat_3d`` / ``mask``:
      split == 'train'  -> rows of the voxels that have a feature, in voxel order (:126-145)
      otherwise         -> one row per voxel, zeros where there is none (:107-111,165-170)
    legacy_mask: the bool form of the old three-key files' 'mask' entry (:112-115,146-149)."""
    dev = torch.device(device)
    feat = torch.as_tensor(feat_3d).to(dev)
    mask = torch.as_tensor(mask_full).to(dev).bool()
    vox = torch.as_tensor(vox_ind).to(dev).long().contiguous()
    if feat.dim() > 2:
        feat = feat[..., 0]                                       # :117-118
    if legacy_mask is not None:
        assert split == 'train', "the three-key format has no val / test branch in the reference"
        lm = torch.as_tensor(legacy_mask).to(dev).bool()
        feat = feat[lm]
        mask = mask.clone()
        mask[mask.clone()] = lm
    feat = feat.contiguous()
    n_pts, n_vox, m_rows = mask.numel(), vox.numel(), feat.shape[0]
    row_bytes = feat.shape[1] * feat.element_size()
    keep_all = int(split != 'train')
    with torch.cuda.device(dev):
        mask_u8 = mask.to(torch.uint8).contiguous()
        mask_vox = torch.empty(n_vox, dtype=torch.uint8, device=dev)
        out = torch.empty((n_vox, feat.shape[1]), dtype=feat.dtype, device=dev)
        ws_bytes = C.lib().osb_feature_remap_workspace_bytes(n_pts, n_vox)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        n_out = ctypes.c_int64(0)
        C.call('osb_feature_remap', C.ptr(mask_u8), n_pts, C.ptr(vox), n_vox, C.ptr(feat), m_rows, row_bytes, keep_all,
               C.ptr(mask_vox), C.ptr(out), ctypes.byref(n_out), C.ptr(ws), ws_bytes, C.stream_ptr())
    return out[:n_out.value], mask_vox.bool()
