"""PyTorch-CPU restatement of the MinkowskiEngine surface used by OpenScene.
TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned** against real
MinkowskiEngine (not available offline); independent of the product's host logic.

Names restated (call sites in the reference):
  SparseTensor(features, coordinates)               run/evaluate.py:284, run/distill.py:316
  MinkowskiConvolution(in, out, kernel_size=, stride=, dilation=, dimension=)
                                                    models/mink_unet.py:47-48,52-53; resnet_base.py:92-97
  MinkowskiConvolutionTranspose(in, out, kernel_size=2, stride=2, dimension=)   mink_unet.py:79-80
  MinkowskiBatchNorm(C) with attribute .bn          mink_unet.py:50; resnet_base.py:79-80
  MinkowskiReLU(inplace=True), cat(a, b)            mink_unet.py:114,147
  modules.resnet_block.{BasicBlock,Bottleneck}      mink_unet.py:26
  utils.kaiming_normal_                             resnet_base.py:76
  MinkowskiAvgPooling / MinkowskiGlobalMaxPooling / MinkowskiLinear   resnet_base.py:54,68,70 (dead code for run/*)

Algorithm (generalised sparse convolution, Choy et al. CVPR'19, cited mink_unet.py:21-23):
  out[o,:] = sum_k sum_{(i,o) in M_k} in[i,:] @ W[k]
with M_k = {(i,o) : c_i = c_o + delta_k * tensor_stride_in * dilation}.  Kernel offsets enumerate
x fastest; odd kernels are centred, even kernels use delta in {0..k-1}.
Everything runs in the dtype of the features (fp64 = truth, fp32 = reference precision).
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn

_R = 1 << 20          # per-axis radix of the packed lookup key
_O = 1 << 19          # offset so negative coordinates pack to non-negative fields


def _pack(c):
    """c: int64 [N,4] (b,x,y,z) -> int64 key, unique per coordinate for |x|<2**19, b<2**3."""
    c = c.astype(np.int64)
    return ((c[:, 0] * _R + (c[:, 1] + _O)) * _R + (c[:, 2] + _O)) * _R + (c[:, 3] + _O)


def kernel_offsets(kernel_size, D=3):
    """Offsets of a hypercube kernel, index k enumerating dimension 0 (x) fastest."""
    if isinstance(kernel_size, int):
        kernel_size = (kernel_size,) * D
    offs = []
    vol = int(np.prod(kernel_size))
    for k in range(vol):
        r, o = k, []
        for d in range(D):
            ks = kernel_size[d]
            idx = r % ks
            r //= ks
            o.append(idx - ks // 2 if ks % 2 == 1 else idx)
        offs.append(o)
    return np.asarray(offs, dtype=np.int64)


class CoordinateManager:
    """Per-tensor-stride coordinate sets and cached kernel maps."""

    def __init__(self, coords):
        c = np.asarray(coords, dtype=np.int64)
        assert c.ndim == 2 and c.shape[1] == 4
        key = _pack(c)
        assert len(np.unique(key)) == len(key), "duplicate coordinates (reference inputs are unique per scene)"
        self.coords = {1: c}
        self._sorted = {}
        self.kmaps = {}

    def _lookup(self, ts, query):
        """rows of ``query`` coordinates in the set at tensor stride ts, -1 where absent."""
        if ts not in self._sorted:
            key = _pack(self.coords[ts])
            order = np.argsort(key, kind='stable')
            self._sorted[ts] = (key[order], order)
        skey, order = self._sorted[ts]
        q = _pack(query)
        pos = np.searchsorted(skey, q)
        pos_c = np.minimum(pos, len(skey) - 1)
        hit = skey[pos_c] == q
        return np.where(hit, order[pos_c], -1)

    def stride(self, ts, s):
        """Coordinates at tensor stride ts*s = unique(floor(c / (ts*s)) * (ts*s))  (SURVEY 8a a5)."""
        new = ts * s
        if new not in self.coords:
            c = self.coords[ts].copy()
            c[:, 1:] = np.floor_divide(c[:, 1:], new) * new
            self.coords[new] = np.unique(c, axis=0)
        return new

    def kernel_map(self, ts_in, ts_out, kernel_size, dilation=1):
        """list over k of (in_rows, out_rows): in = out + delta_k * ts_in * dilation."""
        key = (ts_in, ts_out, kernel_size, dilation)
        if key not in self.kmaps:
            cin, cout = self.coords[ts_in], self.coords[ts_out]
            maps = []
            for d in kernel_offsets(kernel_size):
                q = cout.copy()
                q[:, 1:] += d * ts_in * dilation
                rows = self._lookup(ts_in, q)
                o = np.nonzero(rows >= 0)[0]
                maps.append((torch.from_numpy(rows[o]), torch.from_numpy(o)))
            self.kmaps[key] = maps
        return self.kmaps[key]


class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_manager=None, tensor_stride=1, **kw):
        self.F = features
        if coordinate_manager is None:
            assert coordinates is not None
            c = coordinates.detach().cpu().numpy() if torch.is_tensor(coordinates) else np.asarray(coordinates)
            assert c.shape[0] == features.shape[0]
            coordinate_manager = CoordinateManager(c)
        self.coordinate_manager = coordinate_manager
        self.tensor_stride = tensor_stride

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.coords[self.tensor_stride]).int()

    @property
    def D(self):
        return 3

    def _new(self, F, ts=None):
        return SparseTensor(F, coordinate_manager=self.coordinate_manager,
                            tensor_stride=self.tensor_stride if ts is None else ts)

    def __add__(self, other):
        assert other.tensor_stride == self.tensor_stride
        return self._new(self.F + other.F)

    def __iadd__(self, other):
        assert other.tensor_stride == self.tensor_stride
        self.F = self.F + other.F
        return self


def cat(*tensors):
    ts = tensors[0].tensor_stride
    assert all(t.tensor_stride == ts and t.coordinate_manager is tensors[0].coordinate_manager for t in tensors)
    return tensors[0]._new(torch.cat([t.F for t in tensors], dim=1))


def _conv_apply(F, maps, W, n_out):
    out = F.new_zeros((n_out, W.shape[-1]))
    for k, (ii, oo) in enumerate(maps):
        if len(ii):
            out.index_add_(0, oo, F[ii] @ W[k])
    return out


class _ConvBase(nn.Module):
    TRANSPOSE = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, dimension=None):
        super().__init__()
        assert dimension == 3 and not expand_coordinates and kernel_generator is None
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, dilation, dimension
        self.kernel_volume = kernel_size ** 3
        self.use_mm = (self.kernel_volume == 1 and stride == 1)
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # [ME-upstream, UNVERIFIED]: uniform(-1/sqrt(n), 1/sqrt(n)), n = (out if transpose else in) * volume
        n = (self.out_channels if self.TRANSPOSE else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)


class MinkowskiConvolution(_ConvBase):
    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        if self.use_mm:
            out = x.F @ self.kernel
            ts_out = ts
        else:
            ts_out = cm.stride(ts, self.stride) if self.stride > 1 else ts
            maps = cm.kernel_map(ts, ts_out, self.kernel_size, self.dilation)
            out = _conv_apply(x.F, maps, self.kernel, cm.coords[ts_out].shape[0])
        if self.bias is not None:
            out = out + self.bias
        return x._new(out, ts_out)


class MinkowskiConvolutionTranspose(_ConvBase):
    TRANSPOSE = True

    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        assert ts % self.stride == 0
        ts_out = ts // self.stride
        assert ts_out in cm.coords, "transposed conv needs the cached finer coordinate set (SURVEY 8a a8)"
        # forward map of the matching strided conv (fine -> coarse), used with in/out swapped
        maps = cm.kernel_map(ts_out, ts, self.kernel_size, self.dilation)
        maps_t = [(oo, ii) for (ii, oo) in maps]
        out = _conv_apply(x.F, maps_t, self.kernel, cm.coords[ts_out].shape[0])
        if self.bias is not None:
            out = out + self.bias
        return x._new(out, ts_out)


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return x._new(self.bn(x.F))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._new(torch.relu(x.F))


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return x._new(self.linear(x.F))


class _PoolBase(nn.Module):
    def __init__(self, kernel_size, stride=1, dilation=1, dimension=None, **kw):
        super().__init__()
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation


class MinkowskiAvgPooling(_PoolBase):
    """Average over the inputs present in the kernel region [ME-upstream, UNVERIFIED]."""

    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        ts_out = cm.stride(ts, self.stride) if self.stride > 1 else ts
        maps = cm.kernel_map(ts, ts_out, self.kernel_size, self.dilation)
        n_out = cm.coords[ts_out].shape[0]
        out = x.F.new_zeros((n_out, x.F.shape[1]))
        cnt = x.F.new_zeros((n_out, 1))
        for ii, oo in maps:
            out.index_add_(0, oo, x.F[ii])
            cnt.index_add_(0, oo, x.F.new_ones((len(oo), 1)))
        return x._new(out / cnt.clamp(min=1), ts_out)


class MinkowskiSumPooling(_PoolBase):
    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        ts_out = cm.stride(ts, self.stride) if self.stride > 1 else ts
        maps = cm.kernel_map(ts, ts_out, self.kernel_size, self.dilation)
        out = x.F.new_zeros((cm.coords[ts_out].shape[0], x.F.shape[1]))
        for ii, oo in maps:
            out.index_add_(0, oo, x.F[ii])
        return x._new(out, ts_out)


class MinkowskiGlobalMaxPooling(nn.Module):
    def __init__(self, dimension=None, **kw):
        super().__init__()

    def forward(self, x):
        c = x.coordinate_manager.coords[x.tensor_stride]
        b = torch.from_numpy(c[:, 0])
        nb = int(b.max()) + 1
        out = torch.stack([x.F[b == i].max(0)[0] for i in range(nb)])
        return out          # dense [B, C]; only reachable from dead code in the reference


# ---- MinkowskiEngine.modules.resnet_block -------------------------------------------------------
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation,
                                          dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation,
                                          dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out += residual
        return self.relu(out)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride, dilation=dilation,
                                          dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv3 = MinkowskiConvolution(planes, planes * self.expansion, kernel_size=1, dimension=dimension)
        self.norm3 = MinkowskiBatchNorm(planes * self.expansion, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out += residual
        return self.relu(out)


# ---- MinkowskiEngine.utils ----------------------------------------------------------------------
def _fans(tensor):
    # [ME-upstream, UNVERIFIED]: [vol, in, out] kernels; 2-D kernels follow torch's Linear convention
    if tensor.dim() == 2:
        return tensor.size(1), tensor.size(0)
    rf = tensor.size(0)
    return tensor.size(1) * rf, tensor.size(2) * rf


def kaiming_normal_(tensor, a=0, mode='fan_in', nonlinearity='leaky_relu'):
    fan_in, fan_out = _fans(tensor)
    fan = fan_in if mode == 'fan_in' else fan_out
    gain = nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)


def as_module():
    """This oracle as the ``ME`` namespace argument of ``openscene_b200.minkunet.mink_unet`` / ``synth.build_model``
    (tests, smoke() and bench.py's CPU arm only)."""
    me = sys.modules[__name__]
    return types.SimpleNamespace(**{k: getattr(me, k) for k in dir(me) if not k.startswith('_')})


def install_as_minkowski_engine():
    """Register this oracle as ``MinkowskiEngine`` in sys.modules so that the reference's
    unmodified ``models/*.py`` can be imported on top of it (golden generation only)."""
    me = sys.modules[__name__]
    modules = types.ModuleType('MinkowskiEngine.modules')
    rb = types.ModuleType('MinkowskiEngine.modules.resnet_block')
    rb.BasicBlock, rb.Bottleneck = BasicBlock, Bottleneck
    modules.resnet_block = rb
    utils = types.ModuleType('MinkowskiEngine.utils')
    utils.kaiming_normal_ = kaiming_normal_
    me.modules, me.utils = modules, utils
    sys.modules['MinkowskiEngine'] = me
    sys.modules['MinkowskiEngine.modules'] = modules
    sys.modules['MinkowskiEngine.modules.resnet_block'] = rb
    sys.modules['MinkowskiEngine.utils'] = utils
    return me
