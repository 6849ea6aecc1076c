"""The invariant behind the occupancy-grid lookup (csrc/common.cuh), checked with a NumPy model on the CPU:
rows sorted by the library's Morton key (x least significant, 18-bit biased fields) are contiguous inside every aligned
4x4x4 block and ordered by their in-block Morton index, so  row(cell) = first_row[word] + popcount(word & below)."""
import numpy as np
import pytest

from openscene_b200 import synth
from openscene_b200.coords import _deinterleave


def _spread3(v):
    out = np.zeros_like(v, dtype=np.uint64)
    for i in range(18):
        out |= ((v.astype(np.uint64) >> np.uint64(i)) & np.uint64(1)) << np.uint64(3 * i)
    return out


def _morton(c, bias):
    return _spread3(c[:, 0] + bias) | (_spread3(c[:, 1] + bias) << np.uint64(1)) | (_spread3(c[:, 2] + bias) << np.uint64(2))


@pytest.mark.parametrize('ts', [1, 4])
def test_row_equals_first_row_plus_popcount(ts):
    c = synth.random_cloud(4000, 70, seed=3)[:, 1:].astype(np.int64) * ts          # a set at tensor stride ts
    key = _morton(c, 1 << 17)                                                       # the library's sort key (common.cuh morton_key)
    c = c[np.argsort(key, kind='stable')]
    cell = _morton(c // ts, 0)                                                      # grid cell index (occgrid_cell)
    assert np.all(np.diff(cell.astype(np.int64)) > 0)                               # same order as the rows
    word, bit = cell >> np.uint64(6), (cell & np.uint64(63)).astype(np.int64)
    bitmap, first_row = {}, {}
    for r, (w, b) in enumerate(zip(word.tolist(), bit.tolist())):
        bitmap[w] = bitmap.get(w, 0) | (1 << b)
        first_row.setdefault(w, r)
    for r, (w, b) in enumerate(zip(word.tolist(), bit.tolist())):
        assert first_row[w] + bin(bitmap[w] & ((1 << b) - 1)).count('1') == r


def test_grid_plan_bits_from_key_reductions():
    c = np.array([[0, 5, 300, 7], [2, 0, 1, 129]], dtype=np.int64)
    keys = [(int(b) << 54) | int(_morton(np.array([[x, y, z]]), 1 << 17)[0]) for b, x, y, z in c]
    k_or, k_and = keys[0] | keys[1], keys[0] & keys[1]
    assert all((_deinterleave(k_and, a) >> 17) & 1 for a in range(3))               # every coordinate >= 0
    assert max(_deinterleave(k_or, a) & 0x1ffff for a in range(3)).bit_length() == 9   # 300 < 2^9
    assert (k_or >> 54) + 1 == 3
    neg = (0 << 54) | int(_morton(np.array([[-1, 2, 3]]), 1 << 17)[0])
    assert not all((_deinterleave(neg & k_and, a) >> 17) & 1 for a in range(3))
