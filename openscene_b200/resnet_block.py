"""``MinkowskiEngine.modules.resnet_block`` surface (imported at models/mink_unet.py:26 and
models/resnet_base.py:28): BasicBlock / Bottleneck with the attribute names checkpoints expect
(conv1, norm1, conv2, norm2, [conv3, norm3], relu, downsample)."""
import torch.nn as nn

from .me import MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiReLU


class BasicBlock(nn.Module):
    expansion = 1
    NORM_TYPE = 'BN'

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        conv = lambda i, o, s: MinkowskiConvolution(i, o, kernel_size=3, stride=s, dilation=dilation,
                                                    dimension=dimension)
        self.conv1 = conv(inplanes, planes, stride)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = conv(planes, planes, 1)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.norm2(self.conv2(y))
        y += x if self.downsample is None else self.downsample(x)
        return self.relu(y)


class Bottleneck(nn.Module):
    expansion = 4
    NORM_TYPE = 'BN'

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
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        y = self.norm3(self.conv3(y))
        y += x if self.downsample is None else self.downsample(x)
        return self.relu(y)
