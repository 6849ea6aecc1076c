"""Self-consistency of the CPU oracle: the sparse convolution must equal a dense cross-correlation
evaluated at the active sites (kernel offsets enumerate x fastest; odd kernels centred; even kernels
delta in {0,1}), and strided / transposed maps must be mutual transposes."""
import numpy as np
import torch
import torch.nn.functional as Fn

from oracle import me_cpu
from openscene_b200 import synth
from tests.util import kmap_triples


def _dense(coords, feats, extent):
    g = torch.zeros((1, feats.shape[1], extent, extent, extent), dtype=feats.dtype)   # [1,C,z,y,x]
    g[0, :, coords[:, 3], coords[:, 2], coords[:, 1]] = feats.t()
    return g


def test_conv3_matches_dense_cross_correlation():
    torch.manual_seed(0)
    E = 12
    c = synth.random_cloud(400, E, seed=1)
    f = torch.randn(len(c), 5, dtype=torch.float64)
    conv = me_cpu.MinkowskiConvolution(5, 7, kernel_size=3, dimension=3).double()
    out = conv(me_cpu.SparseTensor(f, torch.from_numpy(c))).F
    # W[k] with k = ix + 3*iy + 9*iz  ->  dense weight [out, in, kz, ky, kx]
    w = conv.kernel.detach().view(3, 3, 3, 5, 7).permute(4, 3, 0, 1, 2).contiguous()
    dense = Fn.conv3d(_dense(c, f, E), w, padding=1)
    ref = dense[0, :, c[:, 3], c[:, 2], c[:, 1]].t()
    assert torch.allclose(out, ref, atol=1e-10)


def test_strided_conv2_even_kernel_offsets():
    torch.manual_seed(0)
    E = 8
    c = synth.random_cloud(200, E, seed=2)
    f = torch.randn(len(c), 3, dtype=torch.float64)
    conv = me_cpu.MinkowskiConvolution(3, 4, kernel_size=2, stride=2, dimension=3).double()
    y = conv(me_cpu.SparseTensor(f, torch.from_numpy(c)))
    assert y.tensor_stride == 2
    cc = y.coordinate_manager.coords[2]
    assert (cc[:, 1:] % 2 == 0).all()
    w = conv.kernel.detach().view(2, 2, 2, 3, 4).permute(4, 3, 0, 1, 2).contiguous()
    dense = Fn.conv3d(_dense(c, f, E), w, stride=2)       # delta in {0,1}: no padding
    ref = dense[0, :, cc[:, 3] // 2, cc[:, 2] // 2, cc[:, 1] // 2].t()
    assert torch.allclose(y.F, ref, atol=1e-10)


def test_transpose_reuses_cached_coordinates_and_is_adjoint():
    torch.manual_seed(0)
    c = synth.random_cloud(300, 10, seed=3, batch=2)
    f = torch.randn(len(c), 4, dtype=torch.float64)
    x = me_cpu.SparseTensor(f, torch.from_numpy(c))
    down = me_cpu.MinkowskiConvolution(4, 4, kernel_size=2, stride=2, dimension=3).double()
    up = me_cpu.MinkowskiConvolutionTranspose(4, 4, kernel_size=2, stride=2, dimension=3).double()
    y = down(x)
    z = up(y)
    assert z.tensor_stride == 1 and z.F.shape[0] == len(c)
    # <down(x), y'> == <x, up_with_same_weights^T(y')>: adjointness of the swapped map
    up.kernel.data = down.kernel.data.transpose(1, 2).contiguous()
    yp = torch.randn_like(y.F)
    lhs = (y.F * yp).sum()
    rhs = (x.F * up(y._new(yp)).F).sum()
    assert torch.allclose(lhs, rhs, atol=1e-9)
    # every fine voxel has exactly one parent pair
    maps = x.coordinate_manager.kernel_map(1, 2, 2, 1)
    assert sum(len(ii) for ii, _ in maps) == len(c)


def test_stride_floor_division_for_negative_coordinates():
    c = np.array([[0, -1, -2, -3], [0, 0, 1, 2], [0, -4, 3, -1]], dtype=np.int64)
    cm = me_cpu.CoordinateManager(c)
    cm.stride(1, 2)
    got = {tuple(r) for r in cm.coords[2].tolist()}
    assert got == {(0, -2, -2, -4), (0, 0, 0, 2), (0, -4, 2, -2)}
