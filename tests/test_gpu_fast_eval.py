"""`model(sinput)` -- the reference's unmodified call (run/evaluate.py:284-289) -- reaches the fused engine through
openscene_b200/fast_eval.py: root discovery, validation against the module path, fallbacks, re-validation after a weight
update.  Results are checked against the golden activations of the reference topology (fp64 oracle)."""
import types

import numpy as np
import pytest
import torch

from openscene_b200 import synth
from tests.util import golden, rel_row_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _input(g):
    import MinkowskiEngine as ME
    return ME.SparseTensor(torch.from_numpy(g['feats']).to(DEV), torch.from_numpy(g['coords']).to(DEV))


def test_unmodified_call_site_reaches_the_engine():
    from openscene_b200 import fast_eval, minkunet
    assert fast_eval.enabled()
    g = golden('unet_MinkUNet18A.npz')
    cfg = types.SimpleNamespace(arch_3d='MinkUNet18A', feature_2d_extractor='openseg')
    torch.manual_seed(0)
    model = minkunet.DisNet(cfg)                                  # DisNet -> net3d (models/disnet.py:21-40)
    synth.randomize_bn_stats(model.net3d, 1)
    model = model.eval().to(DEV)
    outs = []
    with torch.no_grad():                                         # run/evaluate.py:260
        for i in range(4):
            outs.append(model(_input(g)))
    ff = getattr(model.net3d, '_osb_fast', None)
    assert ff is not None and not hasattr(model, '_osb_fast')     # the INNERMOST module that receives the tensor is the root
    assert ff.disabled is None and ff.validated and ff.calls_fast >= 2 and ff.last_err < 1e-3
    for o in outs:
        assert rel_row_err(o.cpu().numpy()[g['rows']], g['out_rows']) < 1e-3
    # training mode / autograd fall back to the module path (BatchNorm statistics, gradients)
    n_fast = ff.calls_fast
    model.train()
    with torch.no_grad():
        model(_input(g))
    model.eval()
    out_grad = model(_input(g))                                   # grad enabled: SparseTensor does not even arm the hook
    assert ff.calls_fast == n_fast and out_grad.requires_grad
    # the train-mode call moved the BatchNorm running statistics: buffers changed -> re-fold + re-validate
    with torch.no_grad():
        b1 = model(_input(g))                                     # validation call (module-path result)
        b2 = model(_input(g))                                     # engine again
    assert ff.disabled is None and ff.validated and ff.calls_fast == n_fast + 1
    assert rel_row_err(b2.cpu().numpy(), b1.cpu().numpy()) < 1e-3
    # a weight update is noticed as well: new results
    with torch.no_grad():
        model.net3d.final.kernel.mul_(2.0)
        o2 = model(_input(g))
        o3 = model(_input(g))
    assert ff.disabled is None and ff.calls_fast == n_fast + 2
    assert rel_row_err(o2.cpu().numpy(), 2.0 * b1.cpu().numpy()) < 1e-3
    assert rel_row_err(o3.cpu().numpy(), 2.0 * b1.cpu().numpy()) < 1e-3


def test_non_minkunet_roots_keep_the_module_path():
    import MinkowskiEngine as ME
    from openscene_b200 import fast_eval

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.c = ME.MinkowskiConvolution(3, 32, kernel_size=3, dimension=3)
            self.r = ME.MinkowskiReLU()

        def forward(self, x):
            return self.r(self.c(x)).F

    g = golden('unet_MinkUNet18A.npz')
    net = Tiny().eval().to(DEV)
    with torch.no_grad():
        a = net(_input(g))
        b = net(_input(g))
        c = net(_input(g))
    assert torch.equal(a, b) and torch.equal(b, c)
    assert net._osb_fast.disabled is not None and 'MinkUNet' in net._osb_fast.disabled
    fast_eval.uninstall(net)
    assert not hasattr(net, '_osb_fast')


def test_lazy_sparse_tensor_surface():
    """the input tensor builds its coordinate manager on first use; `.F` / `.C` / shape need no native work"""
    import MinkowskiEngine as ME
    c = torch.from_numpy(synth.random_cloud(500, 12, seed=3)).to(DEV)
    f = torch.rand(c.shape[0], 3, device=DEV)
    with torch.no_grad():
        x = ME.SparseTensor(f, c)
    assert x._cm is None and x.shape == (c.shape[0], 3) and len(x) == c.shape[0] and x.F is f
    assert torch.equal(x.C, c.int()) and x._cm is not None
    with pytest.raises(ValueError):
        ME.SparseTensor(f[:-1], c)
    dup = torch.cat([c, c[:1]])
    y = ME.SparseTensor(torch.cat([f, f[:1]]), dup)
    with pytest.raises(RuntimeError, match='duplicate'):
        y.coordinate_manager


def test_pipeline_points_to_labels():
    """points -> voxeliser -> network -> labels per point, folded head and materialised head, against the step-by-step path"""
    from openscene_b200 import engine, matching, pipeline
    from openscene_b200.voxelize import voxelize_points
    pts = torch.from_numpy(synth.room_points((0.9, 0.7, 0.6), 2, seed=5)).to(DEV)
    model = synth.build_model('MinkUNet18A', 768, seed=0).eval().to(DEV)
    text = torch.from_numpy(synth.text_embeddings(20)).to(DEV)
    seg = pipeline.OpenVocabSegmenter(model, text, voxel_size=0.02)
    label, scores = seg.segment_points(pts, want_scores=True)
    assert label.shape == (pts.shape[0],) and label.dtype == torch.int64 and scores.shape == (pts.shape[0], 20)
    M = np.eye(4); M[0, 0] = M[1, 1] = M[2, 2] = 1 / 0.02
    cv, inds, inv, _ = voxelize_points(pts, M)
    coords = torch.cat([torch.zeros((cv.shape[0], 1), dtype=torch.int32, device=DEV), cv], 1)
    out = engine.FusedMinkUNet(model)(coords, torch.ones(cv.shape[0], 3, device=DEV))
    s_ref, l_ref, _ = matching._scores(out, inv, text, normalize=True)
    assert (scores.float() - s_ref.float()).abs().max() < 2e-3
    assert (label == l_ref).float().mean() > 0.99
    lab2, _, feat = seg.segment_voxels(coords, torch.ones(cv.shape[0], 3, device=DEV), inv, want_features=True)
    assert torch.equal(lab2, l_ref) and feat.shape == (cv.shape[0], 768)
