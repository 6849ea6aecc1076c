"""Size-independent properties at BASELINE.json's full sizes (the oracle would take minutes there), plus edge cases:
kernel-map symmetry, parent consistency, linearity of the sparse convolution, engine == module surface on the
197k-voxel scene, voxeliser idempotence / inverse, argmax consistency of the matcher, single-voxel and ragged inputs."""
import numpy as np
import pytest
import torch

from openscene_b200 import synth
from tests.util import rel_row_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def scene():
    return torch.from_numpy(synth.scene('config2_200k')).to(DEV)


def test_kernel_map_symmetry_and_counts_at_200k(scene):
    from openscene_b200.coords import CoordinateManager
    cm = CoordinateManager(scene)
    n = cm.sets[1].n
    assert n == scene.shape[0] > 190_000
    km = cm.kernel_map(1, 1, 3)
    nbr = km.nbr
    rows = torch.arange(n, device=DEV, dtype=torch.int32)
    assert torch.equal(nbr[13], rows)                                   # centre offset is the identity
    for k in (0, 5, 12):                                                # nbr[k][o] = i  <=>  nbr[26-k][i] = o
        o = (nbr[k] >= 0).nonzero()[:, 0]
        i = nbr[k][o].long()
        assert torch.equal(nbr[26 - k][i].long(), o)
    assert km.num_pairs() == int((nbr >= 0).sum())
    assert 12.0 < km.num_pairs() / n < 14.5                             # 12-14 pairs per voxel indoors
    ts = 1
    for _ in range(4):                                                  # every fine voxel has exactly one parent
        new = cm.stride(ts, 2)
        par = cm.parent_of[(ts, new)].long()
        fine, coarse = cm.sets[ts].coords.long(), cm.sets[new].coords.long()
        exp = fine.clone()
        exp[:, 1:] = torch.div(fine[:, 1:], new, rounding_mode='floor') * new
        assert torch.equal(coarse[par], exp)
        assert len(torch.unique(par)) == cm.sets[new].n
        down = cm.kernel_map(ts, new, 2)
        assert int((down.nbr >= 0).sum()) == cm.sets[ts].n
        ts = new
    # internal order is a permutation of the caller's order
    assert torch.equal(torch.sort(cm.perm.long())[0], torch.arange(n, device=DEV))
    assert torch.equal(cm.perm[cm.inv_perm.long()].long(), torch.arange(n, device=DEV))


def test_conv_tc_linearity_and_fp32_agreement_at_200k(scene):
    from openscene_b200 import tc
    from openscene_b200.coords import CoordinateManager
    from openscene_b200 import _cabi as C
    cm = CoordinateManager(scene)
    n = cm.sets[1].n
    nbr = cm.kernel_map(1, 1, 3).nbr
    g = torch.Generator(device=DEV).manual_seed(0)
    x, y = torch.randn(n, 96, device=DEV, generator=g), torch.randn(n, 96, device=DEV, generator=g)
    w = torch.randn(27, 96, 96, device=DEV, generator=g) * 0.03
    wp = tc.pack_weights(w)
    f = lambda t: tc.conv_tc(tc.to_split(t), 96, None, 0, nbr, n, 27, wp, 96, out_split=False, out_f32=True)[1]
    fx, fy, fxy = f(x), f(y), f(2.0 * x - 0.5 * y)
    assert rel_row_err((2.0 * fx - 0.5 * fy).cpu().numpy(), fxy.cpu().numpy()) < 1e-4
    ref = torch.empty_like(fx)                                           # exact-fp32 CUDA-core kernel on the same map
    C.call('osb_conv_fwd_f32', C.ptr(x), 96, C.ptr(nbr), n, 27, C.ptr(w), 96, 96, 0, C.ptr(ref), C.stream_ptr())
    assert rel_row_err(fx.cpu().numpy(), ref.cpu().numpy()) < 1e-4


def test_engine_equals_module_surface_at_200k(scene):
    import MinkowskiEngine as ME
    from openscene_b200 import engine, matching
    model = synth.build_model('MinkUNet18A', 768, seed=0).eval().to(DEV)
    feats = torch.ones(scene.shape[0], 3, device=DEV)
    out = engine.FusedMinkUNet(model)(scene, feats)
    with torch.no_grad():
        ref = model(ME.SparseTensor(feats, scene))
    assert out.shape == ref.shape == (scene.shape[0], 768)
    assert rel_row_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-3
    text = torch.from_numpy(synth.text_embeddings(20)).to(DEV)
    s, l, smax = matching._scores(out, None, text, normalize=True, want_smax=True)
    assert torch.equal(l, s.float().max(1)[1]) and torch.allclose(smax, s.float().max(1)[0])
    assert s.float().abs().max() <= 1.001                                # cosine scores


def test_voxeliser_properties_at_full_size():
    from openscene_b200.voxelize import voxelize_points
    pts = torch.from_numpy(synth.room_points((3.2, 2.67, 2.22), 8, seed=1)).to(DEV)
    M = np.eye(4); M[0, 0] = M[1, 1] = M[2, 2] = 1 / 0.02
    cv, inds, inv, _ = voxelize_points(pts, M)
    n_pts, n_vox = pts.shape[0], cv.shape[0]
    assert n_pts > 400_000 and 150_000 < n_vox < 250_000
    assert len(torch.unique(cv, dim=0)) == n_vox                          # voxels are unique
    assert torch.equal(cv[inv][inds], cv)                                # inverse maps each representative to its voxel
    assert bool((inds[inv] <= torch.arange(n_pts, device=DEV)).all())    # representative = first occurrence
    cv2, inds2, inv2, _ = voxelize_points(cv.double(), np.eye(4))        # idempotence: voxelising voxels changes nothing
    assert cv2.shape[0] == n_vox and torch.equal(cv2[inv2], cv)


def test_edge_cases_single_voxel_ragged_batch_and_empty():
    import MinkowskiEngine as ME
    from openscene_b200 import engine
    from openscene_b200.coords import CoordinateManager
    with pytest.raises(RuntimeError, match='empty'):
        CoordinateManager(torch.zeros((0, 4), dtype=torch.int32, device=DEV))
    model = synth.build_model('MinkUNet14A', 64, seed=0).eval().to(DEV)
    eng = engine.FusedMinkUNet(model)
    one = torch.tensor([[0, 5, 6, 7]], dtype=torch.int32, device=DEV)     # a single voxel survives all five levels
    o1 = eng(one, torch.ones(1, 3, device=DEV))
    with torch.no_grad():
        r1 = model(ME.SparseTensor(torch.ones(1, 3, device=DEV), one))
    assert o1.shape == (1, 64) and rel_row_err(o1.cpu().numpy(), r1.cpu().numpy()) < 1e-3
    # ragged batch: scenes of very different size + negative coordinates; each scene must equal its solo run
    a = synth.random_cloud(3000, 40, seed=1)
    b = synth.random_cloud(37, 9, seed=2); b[:, 0] = 1; b[:, 1:] -= 20
    both = torch.from_numpy(np.concatenate([a, b])).to(DEV)
    fo = torch.rand(len(both), 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    ob = eng(both, fo)
    bs = torch.from_numpy(b).to(DEV).clone(); bs[:, 0] = 0
    osolo = eng(bs, fo[len(a):])
    assert rel_row_err(ob[len(a):].cpu().numpy(), osolo.cpu().numpy()) < 1e-3
