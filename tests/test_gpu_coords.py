"""Coordinate sets and kernel maps from libosb200 against the oracle -- exact (integer / index work).
Canonical form: the SET of (k, in_coord, out_coord) triples (row order inside a level is free)."""
import numpy as np
import pytest
import torch

from openscene_b200 import synth
from tests.util import kmap_triples

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device('cuda:0')


def _triples_gpu(km, cin, cout):
    nbr = km.nbr.cpu().numpy()
    cin, cout = cin.cpu().numpy(), cout.cpu().numpy()
    s = set()
    for k in range(km.K):
        o = np.nonzero(nbr[k] >= 0)[0]
        for oo, ii in zip(o, nbr[k][o]):
            s.add((k, tuple(cin[ii]), tuple(cout[oo])))
    return s


def _triples_oracle(maps, cin, cout):
    s = set()
    for k, (ii, oo) in enumerate(maps):
        for i, o in zip(ii.tolist(), oo.tolist()):
            s.add((k, tuple(cin[i]), tuple(cout[o])))
    return s


@pytest.mark.parametrize('case', ['cloud', 'negative', 'room', 'batch2'])
def test_sets_and_maps_match_oracle(case):
    from openscene_b200.coords import CoordinateManager
    from oracle import me_cpu
    if case == 'cloud':
        c = synth.random_cloud(3000, 40, seed=0)
    elif case == 'negative':
        c = synth.random_cloud(2000, 30, seed=1)
        c[:, 1:] -= 17
    elif case == 'room':
        c = synth.scene('tiny')
    else:
        c = synth.random_cloud(1500, 24, seed=2, batch=2)
    cm = CoordinateManager(torch.from_numpy(c).to(_dev()))
    om = me_cpu.CoordinateManager(c)
    # level 0 keeps every row, perm is a permutation
    perm = cm.perm.cpu().numpy()
    assert sorted(perm.tolist()) == list(range(len(c)))
    assert np.array_equal(cm.sets[1].coords.cpu().numpy(), c[perm])
    assert np.array_equal(cm.coords_external(1).cpu().numpy(), c)
    ts = 1
    for _ in range(4):
        new = cm.stride(ts, 2)
        om.stride(ts, 2)
        got = cm.sets[new].coords.cpu().numpy()
        assert {tuple(r) for r in got.tolist()} == {tuple(r) for r in om.coords[new].tolist()}
        assert len(got) == len(om.coords[new])
        # parent_of is consistent with floor division
        par = cm.parent_of[(ts, new)].cpu().numpy()
        fine = cm.sets[ts].coords.cpu().numpy().astype(np.int64)
        exp = fine.copy()
        exp[:, 1:] = np.floor_divide(fine[:, 1:], new) * new
        assert np.array_equal(got[par], exp)
        # 2x2x2 stride-2 map and 3x3x3 map at this level
        for (ti, to, ks) in [(ts, new, 2), (ts, ts, 3)]:
            km = cm.kernel_map(ti, to, ks)
            t_gpu = _triples_gpu(km, cm.sets[ti].coords, cm.sets[to].coords)
            t_ref = _triples_oracle(om.kernel_map(ti, to, ks), om.coords[ti], om.coords[to])
            assert t_gpu == t_ref
            assert km.num_pairs() == len(t_ref)
        ts = new
    km5 = cm.kernel_map(1, 1, 5)
    assert _triples_gpu(km5, cm.sets[1].coords, cm.sets[1].coords) == \
        _triples_oracle(om.kernel_map(1, 1, 5), om.coords[1], om.coords[1])
    # transposed map is the exact swap
    km = cm.kernel_map(1, 2, 2)
    nt = km.transposed().nbr.cpu().numpy()
    n = km.nbr.cpu().numpy()
    for k in range(8):
        o = np.nonzero(n[k] >= 0)[0]
        assert np.array_equal(nt[k][n[k][o]], o)
        assert (nt[k] >= 0).sum() == len(o)
    assert (nt >= 0).sum() == len(c)          # every fine voxel has exactly one parent


def test_duplicates_and_range_fail_loudly():
    from openscene_b200.coords import CoordinateManager
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3]], dtype=torch.int32, device=_dev())
    with pytest.raises(RuntimeError, match='duplicate'):
        CoordinateManager(c)
    c = torch.tensor([[0, 1 << 18, 2, 3]], dtype=torch.int32, device=_dev())
    with pytest.raises(RuntimeError, match='out of range'):
        CoordinateManager(c)


def test_morton_order_is_spatially_compact():
    from openscene_b200.coords import CoordinateManager
    c = synth.scene('tiny')
    cm = CoordinateManager(torch.from_numpy(c).to(_dev()))
    ci = cm.sets[1].coords.cpu().numpy()[:, 1:].astype(np.float64)
    d_sorted = np.linalg.norm(np.diff(ci, axis=0), axis=1).mean()
    d_input = np.linalg.norm(np.diff(c[:, 1:].astype(np.float64), axis=0), axis=1).mean()
    assert d_sorted < 0.25 * d_input


@pytest.mark.parametrize('case', ['room', 'negative_batch'])
def test_pyramid_call_equals_level_by_level_build(case):
    """osb_coordset_pyramid (sort-free power-of-two strides, counts kept on the device) must produce exactly the sets,
    order and parent links of the generic level-by-level path."""
    from openscene_b200.coords import CoordinateManager
    if case == 'room':
        c = synth.scene('tiny')
    else:
        c = synth.random_cloud(2500, 30, seed=4, batch=3)
        c[:, 1:] -= 13
    ct = torch.from_numpy(c).to(_dev())
    a, b = CoordinateManager(ct, pyramid_levels=4), CoordinateManager(ct)
    assert torch.equal(a.perm, b.perm) and torch.equal(a.sets[1].coords, b.sets[1].coords)
    ts = 1
    for _ in range(4):
        new = b.stride(ts, 2)
        assert a.sets[new].n == b.sets[new].n
        assert torch.equal(a.sets[new].coords, b.sets[new].coords)
        assert torch.equal(a.parent_of[(ts, new)], b.parent_of[(ts, new)])
        assert torch.equal(a.kernel_map(new, new, 3).nbr, b.kernel_map(new, new, 3).nbr)
        ts = new


def test_hand_written_sort_handles_many_tiles_and_wide_keys():
    """The LSD radix sort (csrc/sortscan.cuh) behind the Morton ordering: > 100 tiles, three batches, coordinates spread
    over 15 bits per axis (6 radix passes).  The internal order must be the stable Morton order."""
    from openscene_b200.coords import CoordinateManager
    rng = np.random.RandomState(0)
    c = np.unique(np.concatenate([rng.randint(0, 3, (300_000, 1)), rng.randint(-16000, 16000, (300_000, 3))], 1), axis=0)
    c = c[rng.permutation(len(c))].astype(np.int32)
    cm = CoordinateManager(torch.from_numpy(c).to(_dev()), pyramid_levels=2)
    got = cm.sets[1].coords.cpu().numpy().astype(np.int64)

    def spread(v):
        v = (v + (1 << 17)).astype(np.uint64) & np.uint64(0x3FFFF)
        for sh, m in ((32, 0x001F00000000FFFF), (16, 0x001F0000FF0000FF), (8, 0x100F00F00F00F00F), (4, 0x10C30C30C30C30C3), (2, 0x1249249249249249)):
            v = (v | (v << np.uint64(sh))) & np.uint64(m)
        return v
    key = (got[:, 0].astype(np.uint64) << np.uint64(54)) | spread(got[:, 1]) | (spread(got[:, 2]) << np.uint64(1)) | (spread(got[:, 3]) << np.uint64(2))
    assert (np.diff(key.astype(np.float64)) >= 0).all() and len(np.unique(key)) == len(key)
    assert np.array_equal(got, c[cm.perm.cpu().numpy()].astype(np.int64))
    # coarser sets: unique parents, counts equal numpy's
    for l, ts in ((1, 2), (2, 4)):
        exp = np.unique(np.concatenate([c[:, :1], (c[:, 1:] // ts) * ts], 1), axis=0)
        assert cm.sets[ts].n == len(exp)


def _all_maps(cm):
    out = {}
    ts = 1
    for _ in range(4):
        new = cm.stride(ts, 2)
        out[(ts, new, 2)] = cm.kernel_map(ts, new, 2)
        out[(ts, ts, 3)] = cm.kernel_map(ts, ts, 3)
        ts = new
    out[(ts, ts, 3)] = cm.kernel_map(ts, ts, 3)
    out[(1, 1, 5)] = cm.kernel_map(1, 1, 5)
    out[(1, 1, 3, 2)] = cm.kernel_map(1, 1, 3, dilation=2)
    return out


@pytest.mark.parametrize('case', ['room', 'batch3', 'wide', 'negative', 'single'])
@pytest.mark.parametrize('pyramid', [0, 4])
def test_occupancy_grid_maps_equal_hash_maps(case, pyramid, monkeypatch):
    """The occupancy-grid lookup (csrc/common.cuh) and the hash table must give identical kernel maps; sets that the grid
    cannot represent (negative coordinates, > 2^9 cells per axis) fall back to the hash per level."""
    from openscene_b200.coords import CoordinateManager
    if case == 'room':
        c = synth.scene('tiny')
    elif case == 'batch3':
        c = synth.random_cloud(3000, 45, seed=5, batch=3)
    elif case == 'wide':                    # 3000 cells per axis: levels 0-2 on the hash, 3-4 on the grid
        c = synth.random_cloud(6000, 3000, seed=6)
        c = np.concatenate([c, c + np.array([0, 1, 0, 0], dtype=np.int32), c + np.array([0, 0, 2, 1], dtype=np.int32)])
        c = np.unique(c, axis=0).astype(np.int32)
    elif case == 'negative':
        c = synth.random_cloud(2000, 30, seed=7)
        c[:, 1:] -= 9
    else:                                   # nearly dense 4^3 block at the origin: the smallest grid (one word)
        c = synth.random_cloud(60, 4, seed=8)
    ct = torch.from_numpy(c).to(_dev())
    monkeypatch.setenv('OSB_OCCGRID', '1')
    cm_g = CoordinateManager(ct, pyramid_levels=pyramid)
    maps_g = _all_maps(cm_g)
    monkeypatch.setenv('OSB_OCCGRID', '0')
    cm_h = CoordinateManager(ct, pyramid_levels=pyramid)
    maps_h = _all_maps(cm_h)
    assert all(s.grid is None for s in cm_h.sets.values())
    used = {ts: s.grid is not None for ts, s in cm_g.sets.items()}
    if case in ('room', 'batch3', 'single'):
        assert all(used.values())
    elif case == 'wide':
        assert used == {1: False, 2: False, 4: False, 8: True, 16: True}
    else:
        assert not any(used.values())
    if cm_g.grid_status is not None:
        assert int(cm_g.grid_status.item()) == 0
    for key in maps_h:
        assert torch.equal(cm_g.sets[key[0]].coords, cm_h.sets[key[0]].coords)
        assert torch.equal(maps_g[key].nbr, maps_h[key].nbr), key
        assert torch.equal(maps_g[key].pairs_per_k, maps_h[key].pairs_per_k), key
        assert maps_g[key].num_pairs() > 0
