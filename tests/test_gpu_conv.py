"""Sparse convolution forward / backward through the MinkowskiEngine surface (C-ABI underneath)
against the CPU oracle.  fp32 kernels: tolerance 2e-5 relative per row (accumulation order only)."""
import numpy as np
import pytest
import torch

from openscene_b200 import synth
from tests.util import rel_row_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _pair(c, cin, seed=0, dtype=torch.float32):
    from openscene_b200 import me
    from oracle import me_cpu
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(len(c), cin, generator=g)
    xo = me_cpu.SparseTensor(f.to(dtype), torch.from_numpy(c))
    xg = me.SparseTensor(f.to(DEV), torch.from_numpy(c).to(DEV))
    return xo, xg


def _ext_rows(t_gpu, t_oracle):
    """align a coarse-level GPU tensor with the oracle's row order via coordinates."""
    cg = t_gpu.C.cpu().numpy()
    co = t_oracle.coordinate_manager.coords[t_oracle.tensor_stride]
    key = lambda a: (a[:, 0].astype(np.int64) << 60) + ((a[:, 1].astype(np.int64) + 4096) << 40) + \
        ((a[:, 2].astype(np.int64) + 4096) << 20) + (a[:, 3].astype(np.int64) + 4096)
    og, oo = np.argsort(key(cg)), np.argsort(key(co))
    assert np.array_equal(cg[og], co[oo])
    return og, oo


@pytest.mark.parametrize('cin,cout,ks,stride', [(3, 32, 5, 1), (32, 32, 3, 1), (96, 96, 3, 1), (17, 45, 3, 1),
                                                 (32, 64, 2, 2), (96, 768, 1, 1), (64, 64, 3, 2)])
def test_conv_forward_matches_oracle(cin, cout, ks, stride):
    from openscene_b200 import me
    from oracle import me_cpu
    c = synth.random_cloud(2500, 28, seed=3, batch=2)
    xo, xg = _pair(c, cin)
    torch.manual_seed(1)
    co = me_cpu.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dimension=3)
    cg = me.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dimension=3).to(DEV)
    cg.load_state_dict(co.state_dict())
    with torch.no_grad():
        yo, yg = co(xo), cg(xg)
    og, oo = _ext_rows(yg, yo)
    assert rel_row_err(yg.F.cpu().numpy()[og], yo.F.numpy()[oo]) < 2e-5


def test_transpose_conv_and_cat_match_oracle():
    from openscene_b200 import me
    from oracle import me_cpu
    c = synth.scene('tiny')
    xo, xg = _pair(c, 32)
    torch.manual_seed(2)
    mods_o = [me_cpu.MinkowskiConvolution(32, 64, kernel_size=2, stride=2, dimension=3),
              me_cpu.MinkowskiConvolutionTranspose(64, 48, kernel_size=2, stride=2, dimension=3),
              me_cpu.MinkowskiConvolution(80, 96, kernel_size=3, dimension=3)]
    mods_g = [me.MinkowskiConvolution(32, 64, kernel_size=2, stride=2, dimension=3),
              me.MinkowskiConvolutionTranspose(64, 48, kernel_size=2, stride=2, dimension=3),
              me.MinkowskiConvolution(80, 96, kernel_size=3, dimension=3)]
    for a, b in zip(mods_o, mods_g):
        b.load_state_dict(a.state_dict())
        b.to(DEV)
    with torch.no_grad():
        yo = mods_o[2](me_cpu.cat(mods_o[1](mods_o[0](xo)), xo))
        yg = mods_g[2](me.cat(mods_g[1](mods_g[0](xg)), xg))
    # stride-1 rows are in the caller's order on both sides
    assert rel_row_err(yg.F.cpu().numpy(), yo.F.numpy()) < 2e-5


def test_conv_backward_matches_oracle_autograd():
    from openscene_b200 import me
    from oracle import me_cpu
    c = synth.random_cloud(1200, 20, seed=5)
    torch.manual_seed(3)
    f = torch.randn(len(c), 16)
    fo = f.clone().double().requires_grad_(True)
    fg = f.clone().to(DEV).requires_grad_(True)
    net_o = [me_cpu.MinkowskiConvolution(16, 24, kernel_size=3, dimension=3),
             me_cpu.MinkowskiConvolution(24, 24, kernel_size=2, stride=2, dimension=3),
             me_cpu.MinkowskiConvolutionTranspose(24, 8, kernel_size=2, stride=2, dimension=3)]
    net_g = [me.MinkowskiConvolution(16, 24, kernel_size=3, dimension=3),
             me.MinkowskiConvolution(24, 24, kernel_size=2, stride=2, dimension=3),
             me.MinkowskiConvolutionTranspose(24, 8, kernel_size=2, stride=2, dimension=3)]
    for a, b in zip(net_o, net_g):
        b.load_state_dict(a.state_dict())
        a.double()
        b.to(DEV)
    xo = me_cpu.SparseTensor(fo, torch.from_numpy(c))
    xg = me.SparseTensor(fg, torch.from_numpy(c).to(DEV))
    for m in net_o:
        xo = m(xo)
    for m in net_g:
        xg = m(xg)
    w = torch.randn(len(c), 8, generator=torch.Generator().manual_seed(9))
    (xo.F * w.double()).sum().backward()
    (xg.F * w.to(DEV)).sum().backward()
    assert rel_row_err(fg.grad.cpu().numpy(), fo.grad.numpy()) < 5e-5
    for a, b in zip(net_o, net_g):
        ga, gb = a.kernel.grad.numpy(), b.kernel.grad.cpu().numpy()
        assert np.abs(ga - gb).max() / np.abs(ga).max() < 5e-5


@pytest.mark.parametrize('cin,ks', [(3, 5), (1, 3), (4, 5), (2, 2)])
def test_thin_stem_convolution_forward_and_wgrad_match_oracle(cin, ks):
    """cin <= 4, cout == 32 (the 5x5x5 stem in training mode, models/mink_unet.py:47-49) runs on the lane-per-channel kernels of
    csrc/conv_f32.cu: forward and weight gradient against oracle autograd in fp64."""
    from openscene_b200 import me
    from oracle import me_cpu
    c = synth.random_cloud(3000, 30, seed=11, batch=2)
    g = torch.Generator().manual_seed(4)
    f = torch.randn(len(c), cin, generator=g)
    torch.manual_seed(5)
    co = me_cpu.MinkowskiConvolution(cin, 32, kernel_size=ks, dimension=3)
    cg = me.MinkowskiConvolution(cin, 32, kernel_size=ks, dimension=3)
    cg.load_state_dict(co.state_dict())
    co.double(); cg.to(DEV)
    yo = co(me_cpu.SparseTensor(f.double(), torch.from_numpy(c)))
    yg = cg(me.SparseTensor(f.to(DEV), torch.from_numpy(c).to(DEV)))
    assert rel_row_err(yg.F.detach().cpu().numpy(), yo.F.detach().numpy()) < 2e-5
    w = torch.randn(len(c), 32, generator=g)
    (yo.F * w.double()).sum().backward()
    (yg.F * w.to(DEV)).sum().backward()
    ga, gb = co.kernel.grad.numpy(), cg.kernel.grad.cpu().numpy()
    assert ga.shape == gb.shape and np.abs(ga - gb).max() / np.abs(ga).max() < 5e-5


def test_avg_pool_and_global_max():
    from openscene_b200 import me
    from oracle import me_cpu
    c = synth.random_cloud(800, 16, seed=6, batch=2)
    xo, xg = _pair(c, 8)
    po, pg = me_cpu.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3), me.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)
    yo, yg = po(xo), pg(xg)
    og, oo = _ext_rows(yg, yo)
    assert rel_row_err(yg.F.cpu().numpy()[og], yo.F.numpy()[oo]) < 1e-5
    go, gg = me_cpu.MinkowskiGlobalMaxPooling(dimension=3)(xo), me.MinkowskiGlobalMaxPooling(dimension=3)(xg)
    assert torch.allclose(go, gg.cpu(), atol=1e-6)
