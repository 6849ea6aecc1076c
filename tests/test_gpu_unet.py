"""MinkUNet18A / 34C forward on the GPU (drop-in MinkowskiEngine surface -> C-ABI) against
(a) the golden activations produced by the reference's models/mink_unet.py on the fp64 oracle and
(b) the fp32 oracle run live.  Tolerance (north star): 1e-3 relative per point feature."""
import numpy as np
import pytest
import torch

from openscene_b200 import synth
from tests.util import golden, rel_row_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-3


@pytest.mark.parametrize('arch', ['MinkUNet18A', 'MinkUNet34C'])
def test_forward_matches_reference_golden(arch):
    import MinkowskiEngine as ME          # the drop-in package of this repository
    g = golden(f'unet_{arch}.npz')
    model = synth.build_model(arch, 768, seed=0).eval().to(DEV)
    assert list(model.state_dict().keys()) == g['state_keys'].tolist()
    with torch.no_grad():
        x = ME.SparseTensor(torch.from_numpy(g['feats']).to(DEV), torch.from_numpy(g['coords']).to(DEV))
        out = model(x).cpu().numpy()
    assert out.shape == (len(g['coords']), 768)
    assert rel_row_err(out[g['rows']], g['out_rows']) < TOL
    assert np.allclose(np.linalg.norm(out, axis=1), g['row_norm'], rtol=TOL)


def test_train_mode_batchnorm_and_backward_match_oracle():
    """distill.py runs the net in train mode (BN batch statistics) and back-propagates a cosine loss."""
    from openscene_b200 import me, minkunet
    from oracle import matching as omatch
    from oracle import me_cpu
    c = synth.random_cloud(1500, 20, seed=8, batch=2)
    f = torch.rand(len(c), 3, generator=torch.Generator().manual_seed(0))
    tgt = torch.randn(len(c), 64, generator=torch.Generator().manual_seed(1))
    mo = synth.build_model('MinkUNet14A', 64, seed=0, ME=minkunet.oracle_me()).double().train()
    mg = synth.build_model('MinkUNet14A', 64, seed=0).to(DEV).train()
    oo = mo(me_cpu.SparseTensor(f.double(), torch.from_numpy(c)))
    og = mg(me.SparseTensor(f.to(DEV), torch.from_numpy(c).to(DEV)))
    assert rel_row_err(og.detach().cpu().numpy(), oo.detach().numpy()) < TOL
    lo = omatch.distill_loss(oo, tgt.double())
    lg = (1 - torch.nn.CosineSimilarity()(og, tgt.to(DEV))).mean()
    assert abs(lo.item() - lg.item()) < 1e-5
    lo.backward()
    lg.backward()
    for (n, po), (_, pg) in zip(mo.named_parameters(), mg.named_parameters()):
        a, b = po.grad.numpy(), pg.grad.cpu().numpy()
        assert np.abs(a - b).max() <= 1e-2 * np.abs(a).max() + 1e-9, n   # fp32 GPU vs fp64 oracle through train-mode BN
    # running statistics were updated identically
    assert torch.allclose(mo.bn0.bn.running_mean.float(), mg.bn0.bn.running_mean.cpu(), atol=1e-5)
