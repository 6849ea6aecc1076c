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
    """distill.py runs the net in train mode (BN batch statistics) and back-propagates a cosine loss.
    Truth = fp64 oracle.  Train-mode BN over the handful of voxels of the coarse levels is ill-conditioned, so the
    yardstick for the gradients is the fp64 oracle itself with every conv kernel perturbed by 2^-16 relative noise
    (the operand precision of the bf16x3 tensor-core path): the GPU error must stay within a small multiple of the
    gradient change that perturbation causes.  (The sparse-conv gradients themselves are checked tightly in
    tests/test_gpu_conv.py and tests/test_gpu_conv_tc.py.)"""
    import copy
    from openscene_b200 import me, minkunet
    from oracle import matching as omatch
    from oracle import me_cpu
    c = synth.scene('tiny')
    f = torch.rand(len(c), 3, generator=torch.Generator().manual_seed(0))
    tgt = torch.randn(len(c), 64, generator=torch.Generator().manual_seed(1))
    m64 = synth.build_model('MinkUNet14A', 64, seed=0, ME=me_cpu.as_module()).double().train()
    mpt = copy.deepcopy(m64)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n_, p_ in mpt.named_parameters():
            if n_.endswith('kernel'):
                p_.mul_(1 + 2.0 ** -16 * torch.randn(p_.shape, generator=g, dtype=torch.float64))
    mg = synth.build_model('MinkUNet14A', 64, seed=0).to(DEV).train()
    o64 = m64(me_cpu.SparseTensor(f.double(), torch.from_numpy(c)))
    opt = mpt(me_cpu.SparseTensor(f.double(), torch.from_numpy(c)))
    og = mg(me.SparseTensor(f.to(DEV), torch.from_numpy(c).to(DEV)))
    e_fwd = rel_row_err(og.detach().cpu().numpy(), o64.detach().numpy())
    e_fwd_pert = rel_row_err(opt.detach().numpy(), o64.detach().numpy())
    print('forward rel err: gpu', e_fwd, ' 2^-16-perturbed oracle', e_fwd_pert)
    assert e_fwd < max(TOL, 8 * e_fwd_pert)
    l64 = omatch.distill_loss(o64, tgt.double())
    lpt = omatch.distill_loss(opt, tgt.double())
    lg = (1 - torch.nn.CosineSimilarity()(og, tgt.to(DEV))).mean()
    assert abs(l64.item() - lg.item()) < max(1e-5, 8 * abs(l64.item() - lpt.item()))
    l64.backward(); lpt.backward(); lg.backward()
    worst = 0.0
    for (n, p64), (_, ppt), (_, pg) in zip(m64.named_parameters(), mpt.named_parameters(), mg.named_parameters()):
        a = p64.grad.numpy()
        e_pert = np.abs(a - ppt.grad.numpy()).max()
        e_gpu = np.abs(a - pg.grad.cpu().numpy().astype(np.float64)).max()
        scale = np.abs(a).max()
        worst = max(worst, e_gpu / (e_pert + 1e-4 * scale))
        assert e_gpu <= 10 * e_pert + 1e-3 * scale, (n, e_gpu, e_pert, scale)
    print('worst gpu / perturbation gradient-error ratio:', worst)
    # running statistics were updated identically
    assert torch.allclose(m64.bn0.bn.running_mean.float(), mg.bn0.bn.running_mean.cpu(), atol=1e-5)


def test_distill_step_reduces_loss():
    """run/distill.py:311-334 on the drop-in surface: a few Adam steps on one scene lower the cosine loss."""
    from openscene_b200 import distill
    c = torch.from_numpy(synth.random_cloud(1200, 18, seed=9))
    f = torch.ones(len(c), 3)
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(len(c), generator=g) < 0.6
    tgt = torch.randn(int(mask.sum()), 64, generator=g).half()
    model = synth.build_model('MinkUNet14A', 64, seed=1).to(DEV).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = [float(distill.distill_step(model, opt, c, f, tgt, mask)) for _ in range(6)]
    assert losses[-1] < losses[0] - 0.02, losses


def test_distill_late_head_is_the_same_step():
    """distill_step(late_head=True) applies the final 1x1x1 convolution to the supervised rows only: the loss and every
    parameter gradient must equal the plain ``model(sinput)[mask]`` step (run/distill.py:321-333)."""
    from openscene_b200 import distill
    c = torch.from_numpy(synth.random_cloud(2000, 22, seed=12))
    f = torch.ones(len(c), 3)
    g = torch.Generator().manual_seed(8)
    mask = torch.rand(len(c), generator=g) < 0.1
    tgt = torch.randn(int(mask.sum()), 64, generator=g).half()

    class Keep(torch.optim.SGD):                       # an optimiser that leaves the weights alone: gradients stay comparable
        def step(self, closure=None):
            return None
    grads, losses = [], []
    for late in (False, True):
        model = synth.build_model('MinkUNet14A', 64, seed=1).to(DEV).train()
        opt = Keep(model.parameters(), lr=0.0)
        losses.append(float(distill.distill_step(model, opt, c, f, tgt, mask, translate=False, late_head=late)))
        assert type(model.final).__name__ == 'MinkowskiConvolution'            # restored
        grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
    assert abs(losses[0] - losses[1]) < 1e-5
    for k in grads[0]:
        a, b = grads[0][k], grads[1][k]
        assert (a - b).abs().max() <= 2e-3 * a.abs().max() + 1e-8, k
