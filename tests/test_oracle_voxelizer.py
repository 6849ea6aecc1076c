"""The NumPy restatement of the voxeliser against vectors produced by the REFERENCE's own
dataset/voxelizer.py (scripts/make_golden.py).  Bit-exact: integer / index work."""
import numpy as np
import pytest

from oracle import voxelize_ref
from tests.util import golden

CASES = ['aug_f64', 'noaug_f32', 'dups_f64', 'neg_f64']


@pytest.mark.parametrize('case', CASES)
def test_restatement_matches_reference(case):
    g = golden(f'voxelizer_{case}.npz')
    cv, inds, inv, _ = voxelize_ref.voxelize(g['points'], g['matrix'])
    assert np.array_equal(cv, g['coords_vox'])
    assert np.array_equal(inds, g['inds'])
    assert np.array_equal(inv, g['inds_reverse'])


def test_fnv_wraps_and_orders_like_numpy_unique():
    # voxelization_utils.py:9-22: multiply THEN xor on uint64 words, wrap mod 2**64
    a = np.array([[0, 0, 0], [1, 2, 3], [70000, 5, 9]], dtype=np.float64)
    h = voxelize_ref.fnv_hash_vec(a)
    ref = []
    for row in a.astype(np.uint64).tolist():
        x = 14695981039346656037
        for v in row:
            x = (x * 1099511628211) % (1 << 64)
            x ^= v
        ref.append(x)
    assert h.tolist() == ref


def test_first_occurrence_and_inverse():
    c = np.array([[1, 1, 1], [0, 0, 0], [1, 1, 1], [0, 0, 0], [2, 0, 0]], dtype=np.float64)
    inds, inv = voxelize_ref.sparse_quantize_index(c)
    assert sorted(inds.tolist()) == [0, 1, 4]           # first occurrences
    assert np.array_equal(c[inds][inv], c)              # inverse reconstructs every point's voxel


def test_seeded_matrix_is_rigid_scale():
    rng = np.random.RandomState(0)
    M = voxelize_ref.transformation_matrix(0.02, rng)
    R = M[:3, :3]
    s = np.cbrt(np.linalg.det(R))
    assert 45.0 <= s <= 55.0                             # (1/0.02) * U(0.9,1.1)
    assert np.allclose(R @ R.T, s * s * np.eye(3), atol=1e-9)
