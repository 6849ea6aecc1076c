"""NumPy restatement of the reference voxeliser.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * ``dataset/voxelization_utils.py:9-22``   (fnv_hash_vec: multiply-then-xor on uint64 words)
  * ``dataset/voxelization_utils.py:44-137`` (sparse_quantize, hash_type='fnv', return_index)
  * ``dataset/voxelizer.py:97-140``          (Voxelizer.voxelize with clip_bound=None)

The random part of the reference (``get_transformation_matrix``, voxelizer.py:46-76) is
host-side 4x4 matrix algebra and stays in Python in the product too; here the matrix is an
explicit argument so that the reference, the oracle and the CUDA path can be fed the same one.
"""
import numpy as np

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)


def fnv_hash_vec(arr):
    """voxelization_utils.py:9-22.  arr: [N, D] integral-valued float/int array."""
    assert arr.ndim == 2
    a = arr.astype(np.uint64)                      # float -> uint64 (values are >= 0 here)
    h = np.full(a.shape[0], FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over='ignore'):
        for j in range(a.shape[1]):
            h = h * FNV_PRIME                       # wraps mod 2**64
            h = np.bitwise_xor(h, a[:, j])
    return h


def sparse_quantize_index(discrete_coords):
    """voxelization_utils.py:107-131 with return_index=True, hash_type='fnv', quantization_size=1.

    Returns (inds, inds_reverse): first-occurrence index per distinct key in ascending-key
    order, and the voxel row of every point."""
    key = fnv_hash_vec(np.floor(discrete_coords))
    _, inds, inds_reverse = np.unique(key, return_index=True, return_inverse=True)
    return inds, inds_reverse


def voxelize(coords, rigid_transformation):
    """voxelizer.py:116-140 for a given 4x4 ``rigid_transformation`` (= M_r @ M_v).

    coords: float [N,3].  Returns (coords_vox float64 [Nv,3], inds int64 [Nv],
    inds_reconstruct int64 [N], min_coords float64 [3])."""
    assert coords.shape[1] == 3 and coords.shape[0]
    homo = np.hstack((coords, np.ones((coords.shape[0], 1), dtype=coords.dtype)))
    coords_aug = np.floor(homo @ rigid_transformation.T[:, :3])
    min_coords = coords_aug.min(0)
    coords_aug = np.floor(coords_aug - min_coords)
    inds, inds_reconstruct = sparse_quantize_index(coords_aug)
    return coords_aug[inds], inds.astype(np.int64), np.asarray(inds_reconstruct).astype(np.int64), min_coords


def transformation_matrix(voxel_size, rng, use_augmentation=True,
                          scale_bound=(0.9, 1.1),
                          rot_bound=((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64), (-np.pi, np.pi))):
    """Seedable restatement of voxelizer.py:46-76 + the product M_r @ M_v of :117-121.

    ``rng`` is a ``np.random.RandomState``; draws happen in the reference's order
    (three rotation angles, the shuffle, then the scale)."""
    from scipy.linalg import expm, norm
    M_v, M_r = np.eye(4), np.eye(4)
    rot = np.eye(3)
    if use_augmentation and rot_bound is not None:
        mats = []
        for axis_ind, b in enumerate(rot_bound):
            theta = 0
            axis = np.zeros(3)
            axis[axis_ind] = 1
            if b is not None:
                theta = rng.uniform(*b)
            mats.append(expm(np.cross(np.eye(3), axis / norm(axis) * theta)))
        rng.shuffle(mats)
        rot = mats[0] @ mats[1] @ mats[2]
    M_r[:3, :3] = rot
    scale = 1 / voxel_size
    if use_augmentation and scale_bound is not None:
        scale *= rng.uniform(*scale_bound)
    np.fill_diagonal(M_v[:3, :3], scale)
    return (M_r @ M_v) if use_augmentation else M_v
