"""GPU voxeliser with the reference's interface (``dataset/voxelizer.py``): same constructor arguments,
``get_transformation_matrix`` (host-side 4x4 algebra, NumPy RNG draws in the reference's order) and
``voxelize`` whose transform / floor / FNV hash / unique run on the device (csrc/voxelize.cu) and return
the same ``inds`` / ``inds_reconstruct`` as ``np.unique`` would (ascending FNV key, first occurrence)."""
import collections.abc
import ctypes

import numpy as np
import torch
from scipy.linalg import expm, norm

from . import _cabi as C


def _rot(axis, theta):
    return expm(np.cross(np.eye(3), axis / norm(axis) * theta))


def voxelize_points(points, matrix):
    """points: CUDA float32/float64 [N,3]; matrix: 4x4 float64 (host).  Returns device tensors
    (coords_vox int32 [Nv,3], inds int64 [Nv], inds_reverse int64 [N]) and the subtracted minimum."""
    C.require_cuda(points, 'points')
    assert points.dim() == 2 and points.shape[1] == 3 and points.shape[0] > 0
    if points.dtype not in (torch.float32, torch.float64):
        points = points.double()
    points = points.contiguous()
    n, dev = points.shape[0], points.device
    M = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64))
    assert M.shape == (4, 4)
    with torch.cuda.device(dev):
        ws_bytes = C.lib().osb_voxelize_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cv = torch.empty((n, 3), dtype=torch.int32, device=dev)
        inds = torch.empty(n, dtype=torch.int64, device=dev)
        inv = torch.empty(n, dtype=torch.int64, device=dev)
        nv = ctypes.c_int64(0)
        mn = (ctypes.c_double * 3)()
        C.call('osb_voxelize', C.ptr(points), int(points.dtype == torch.float64), n,
               M.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), C.ptr(cv), C.ptr(inds), C.ptr(inv),
               ctypes.byref(nv), mn, C.ptr(ws), ws_bytes, C.stream_ptr())
    return cv[:nv.value], inds[:nv.value], inv, np.array(list(mn))


class Voxelizer:
    def __init__(self, voxel_size=1, clip_bound=None, use_augmentation=False, scale_augmentation_bound=None,
                 rotation_augmentation_bound=None, translation_augmentation_ratio_bound=None, ignore_label=255):
        if clip_bound is not None:
            raise NotImplementedError("clip_bound is None in every OpenScene loader (dataset/point_loader.py:93-99)")
        self.voxel_size = voxel_size
        self.clip_bound = clip_bound
        self.ignore_label = ignore_label
        self.use_augmentation = use_augmentation
        self.scale_augmentation_bound = scale_augmentation_bound
        self.rotation_augmentation_bound = rotation_augmentation_bound
        self.translation_augmentation_ratio_bound = translation_augmentation_ratio_bound

    def get_transformation_matrix(self):
        """voxelizer.py:46-76 -- returns (voxelization_matrix, rotation_matrix)."""
        M_v, M_r = np.eye(4), np.eye(4)
        rot = np.eye(3)
        if self.use_augmentation and self.rotation_augmentation_bound is not None:
            if not isinstance(self.rotation_augmentation_bound, collections.abc.Iterable):
                raise ValueError()
            mats = []
            for axis_ind, bound in enumerate(self.rotation_augmentation_bound):
                theta, axis = 0, np.zeros(3)
                axis[axis_ind] = 1
                if bound is not None:
                    theta = np.random.uniform(*bound)
                mats.append(_rot(axis, theta))
            np.random.shuffle(mats)
            rot = mats[0] @ mats[1] @ mats[2]
        M_r[:3, :3] = rot
        scale = 1 / self.voxel_size
        if self.use_augmentation and self.scale_augmentation_bound is not None:
            scale *= np.random.uniform(*self.scale_augmentation_bound)
        np.fill_diagonal(M_v[:3, :3], scale)
        return M_v, M_r

    def voxelize(self, coords, feats, labels, center=None, link=None, return_ind=False, device='cuda'):
        """voxelizer.py:97-140.  ``coords`` / ``feats`` / ``labels`` may be NumPy arrays (as in the reference's
        loaders) or torch tensors; results come back as NumPy arrays like the reference's."""
        assert coords.shape[1] == 3 and coords.shape[0] == feats.shape[0] and coords.shape[0]
        M_v, M_r = self.get_transformation_matrix()
        rigid = (M_r @ M_v) if self.use_augmentation else M_v
        pts = torch.as_tensor(coords).to(device)
        cv, inds, inv, _ = voxelize_points(pts, rigid)
        inds_np = inds.cpu().numpy()
        coords_aug = cv.cpu().numpy().astype(np.float64)
        feats, labels = feats[inds_np], labels[inds_np]
        if feats.shape[1] > 6:
            feats[:, 3:6] = feats[:, 3:6] @ (M_r[:3, :3].T)
        inds_rec = inv.cpu().numpy()
        if return_ind:
            return coords_aug, feats, labels, inds_rec, inds_np
        if link is not None:
            return coords_aug, feats, labels, inds_rec, link[inds_np]
        return coords_aug, feats, labels, inds_rec
