"""Table-driven mirror of the reference's MinkUNet family and DisNet wrapper.

The reference builds these nets in ``models/mink_unet.py:44-114`` (layers), ``:116-174`` (forward),
``:176-238`` (variants) on top of ``models/resnet_base.py:73-118`` (init, ``_make_layer``), and wraps
them in ``models/disnet.py:21-40`` (attribute ``net3d``).  The reference tree is not present on the
benchmark machine, so this file re-creates the same module tree from a small spec: identical
attribute names (-> identical state-dict keys: ``conv0p1s1.kernel``, ``bn0.bn.weight``,
``block1.0.conv1.kernel``, ``block2.0.downsample.0.kernel``, ``convtr4p16s2.kernel``, ``final.kernel``),
identical construction order (-> identical seeded initialisation), identical forward dataflow.
``tests/test_topology.py`` checks that against the reference's own file when it is available.

``ME`` is the namespace providing the MinkowskiEngine surface: the product (``openscene_b200.me`` +
``resnet_block``) by default; tests pass the CPU oracle to run the same topology on it.
"""
import types

import torch.nn as nn

# arch -> (block, layers per stage, planes per stage); mink_unet.py:176-238
_L14, _L18, _L34 = (1,) * 8, (2,) * 8, (2, 3, 4, 6, 2, 2, 2, 2)
ARCHS = {
    'MinkUNet14A': ('basic', _L14, (32, 64, 128, 256, 128, 128, 96, 96)),
    'MinkUNet14B': ('basic', _L14, (32, 64, 128, 256, 128, 128, 128, 128)),
    'MinkUNet14C': ('basic', _L14, (32, 64, 128, 256, 192, 192, 128, 128)),
    'MinkUNet14D': ('basic', _L14, (32, 64, 128, 256, 384, 384, 384, 384)),
    'MinkUNet18A': ('basic', _L18, (32, 64, 128, 256, 128, 128, 96, 96)),
    'MinkUNet18B': ('basic', _L18, (32, 64, 128, 256, 128, 128, 128, 128)),
    'MinkUNet18D': ('basic', _L18, (32, 64, 128, 256, 384, 384, 384, 384)),
    'MinkUNet34A': ('basic', _L34, (32, 64, 128, 256, 256, 128, 64, 64)),
    'MinkUNet34B': ('basic', _L34, (32, 64, 128, 256, 256, 128, 64, 32)),
    'MinkUNet34C': ('basic', _L34, (32, 64, 128, 256, 256, 128, 96, 96)),
}
INIT_DIM = 32


def default_me():
    from . import me, me_utils, resnet_block
    ns = types.SimpleNamespace(**{k: getattr(me, k) for k in dir(me) if not k.startswith('_')})
    ns.BasicBlock, ns.Bottleneck = resnet_block.BasicBlock, resnet_block.Bottleneck
    ns.kaiming_normal_ = me_utils.kaiming_normal_
    return ns


class MinkUNet(nn.Module):
    def __init__(self, arch='MinkUNet18A', in_channels=3, out_channels=20, D=3, ME=None):
        super().__init__()
        if arch not in ARCHS:
            raise Exception('architecture not supported yet')
        ME = ME or default_me()
        self._ME = types.SimpleNamespace(cat=ME.cat)
        block_kind, layers, planes = ARCHS[arch]
        block = ME.BasicBlock if block_kind == 'basic' else ME.Bottleneck
        exp = block.expansion
        self.arch, self.D = arch, D
        width = INIT_DIM                       # running "inplanes"

        def stage(planes_i, n_blocks):
            nonlocal width
            down = None
            if width != planes_i * exp:
                down = nn.Sequential(ME.MinkowskiConvolution(width, planes_i * exp, kernel_size=1, stride=1, dimension=D),
                                     ME.MinkowskiBatchNorm(planes_i * exp))
            blocks = [block(width, planes_i, stride=1, dilation=1, downsample=down, dimension=D)]
            width = planes_i * exp
            blocks += [block(width, planes_i, stride=1, dilation=1, dimension=D) for _ in range(1, n_blocks)]
            return nn.Sequential(*blocks)

        self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, width, kernel_size=5, dimension=D)
        self.bn0 = ME.MinkowskiBatchNorm(width)
        # encoder: conv{i}p{2^(i-1)}s2 / bn{i} / block{i}
        skip_width = [width]
        for i in range(1, 5):
            setattr(self, f'conv{i}p{2 ** (i - 1)}s2',
                    ME.MinkowskiConvolution(width, width, kernel_size=2, stride=2, dimension=D))
            setattr(self, f'bn{i}', ME.MinkowskiBatchNorm(width))
            setattr(self, f'block{i}', stage(planes[i - 1], layers[i - 1]))
            skip_width.append(width)
        # decoder: convtr{j}p{2^(8-j)}s2 / bntr{j} / block{j+1}; skip from encoder level 7-j
        for j in range(4, 8):
            setattr(self, f'convtr{j}p{2 ** (8 - j)}s2',
                    ME.MinkowskiConvolutionTranspose(width, planes[j], kernel_size=2, stride=2, dimension=D))
            setattr(self, f'bntr{j}', ME.MinkowskiBatchNorm(planes[j]))
            width = planes[j] + skip_width[7 - j]
            setattr(self, f'block{j + 1}', stage(planes[j], layers[j]))
        self.final = ME.MinkowskiConvolution(planes[7], out_channels, kernel_size=1, dimension=D)
        self.relu = ME.MinkowskiReLU(inplace=True)

        # resnet_base.py:73-80
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.kaiming_normal_(m.kernel, mode='fan_out', nonlinearity='relu')
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def forward(self, x):
        relu, cat = self.relu, self._ME.cat
        skips = [relu(self.bn0(self.conv0p1s1(x)))]
        out = skips[0]
        for i in range(1, 5):
            out = relu(getattr(self, f'bn{i}')(getattr(self, f'conv{i}p{2 ** (i - 1)}s2')(out)))
            out = getattr(self, f'block{i}')(out)
            skips.append(out)
        for j in range(4, 8):
            out = relu(getattr(self, f'bntr{j}')(getattr(self, f'convtr{j}p{2 ** (8 - j)}s2')(out)))
            out = getattr(self, f'block{j + 1}')(cat(out, skips[7 - j]))
        return self.final(out).F


def mink_unet(in_channels=3, out_channels=20, D=3, arch='MinkUNet18A', ME=None):
    """Factory with the reference's signature (models/mink_unet.py:241-263)."""
    return MinkUNet(arch, in_channels, out_channels, D, ME=ME)


class DisNet(nn.Module):
    """models/disnet.py:21-40: 3-D sparse U-Net for distillation; state-dict prefix ``net3d.``."""

    def __init__(self, cfg=None, ME=None):
        super().__init__()
        if not hasattr(cfg, 'feature_2d_extractor'):
            cfg.feature_2d_extractor = 'openseg'
        if 'lseg' in cfg.feature_2d_extractor:
            last_dim = 512
        elif 'openseg' in cfg.feature_2d_extractor:
            last_dim = 768
        else:
            raise NotImplementedError
        self.net3d = mink_unet(in_channels=3, out_channels=last_dim, D=3, arch=cfg.arch_3d, ME=ME)

    def forward(self, sparse_3d):
        return self.net3d(sparse_3d)
