"""Parameter containers with torchvision's ResNet-50 state_dict layout (no torchvision import, no forward math).

The reference builds its encoders from `torchvision.models.resnet50` (src/vision_models/moco.py:11,34,46,78,90);
checkpoints and `EmbeddingNet.state_dict()` therefore use torchvision's key names (`conv1.weight`,
`layer1.0.bn2.running_var`, `layer3.0.5.conv3.weight`, `layer3.1.downsample.0.bias`, ...). These holders reproduce
that naming so checkpoints interchange; the arithmetic lives in the CUDA program (program.py).
"""
import math

import torch
from torch import nn


class ConvP(nn.Module):
    def __init__(self, c_in, c_out, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c_out, c_in, k, k))
        # torchvision/models/resnet.py:208-210: kaiming_normal_(mode="fan_out", nonlinearity="relu")
        nn.init.normal_(self.weight, 0.0, math.sqrt(2.0 / (c_out * k * k)))
        if bias:
            # nn.Conv2d default bias init: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(c_in * k * k)
            self.bias = nn.Parameter(torch.empty(c_out).uniform_(-bound, bound))


class BNP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class BottleneckP(nn.Module):
    expansion = 4

    def __init__(self, c_in, planes, downsample):
        super().__init__()
        self.conv1, self.bn1 = ConvP(c_in, planes, 1), BNP(planes)
        self.conv2, self.bn2 = ConvP(planes, planes, 3), BNP(planes)
        self.conv3, self.bn3 = ConvP(planes, planes * 4, 1), BNP(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(ConvP(c_in, planes * 4, 1), BNP(planes * 4))


class BasicBlockP(nn.Module):
    """torchvision BasicBlock(C, c, downsample=Sequential(Conv2d(C, c, 3, padding=1) [bias=True], BN(c)))."""

    def __init__(self, c_in, planes):
        super().__init__()
        self.conv1, self.bn1 = ConvP(c_in, planes, 3), BNP(planes)
        self.conv2, self.bn2 = ConvP(planes, planes, 3), BNP(planes)
        self.downsample = nn.Sequential(ConvP(c_in, planes, 3, bias=True), BNP(planes))


def _make_layer(c_in, planes, blocks):
    layers = [BottleneckP(c_in, planes, True)]
    layers += [BottleneckP(planes * 4, planes, False) for _ in range(blocks - 1)]
    return nn.Sequential(*layers)


class ResNet50Params(nn.Module):
    """variant: 'conv5' | 'l4' | 'l3' (moco_conv5 / moco_conv4_compressed / moco_conv3_compressed)."""

    OUT = {"conv5": 2048, "l4": 42 * 7 * 7, "l3": 11 * 14 * 14}

    def __init__(self, variant="conv5"):
        super().__init__()
        assert variant in self.OUT
        self.variant = variant
        self.conv1, self.bn1 = ConvP(3, 64, 7), BNP(64)
        self.layer1 = _make_layer(64, 64, 3)
        self.layer2 = _make_layer(256, 128, 4)
        layer3 = _make_layer(512, 256, 6)
        layer4 = _make_layer(1024, 512, 3)
        if variant == "l3":
            self.layer3 = nn.Sequential(layer3, BasicBlockP(1024, 11))
            self.layer4 = nn.Sequential()
        elif variant == "l4":
            self.layer3 = layer3
            self.layer4 = nn.Sequential(layer4, BasicBlockP(2048, 42))
        else:
            self.layer3, self.layer4 = layer3, layer4
        # zero-init is NOT used by torchvision's default constructor (zero_init_residual=False)
        self.out_size = self.OUT[variant]


class BasicP(nn.Module):
    """torchvision BasicBlock (tv:models/resnet.py:59-101): two 3x3 convs, optional 1x1 projection shortcut."""
    expansion = 1

    def __init__(self, c_in, planes, downsample):
        super().__init__()
        self.conv1, self.bn1 = ConvP(c_in, planes, 3), BNP(planes)
        self.conv2, self.bn2 = ConvP(planes, planes, 3), BNP(planes)
        if downsample:
            self.downsample = nn.Sequential(ConvP(c_in, planes, 1), BNP(planes))


class ResNetBasicParams(nn.Module):
    """resnet18 / resnet34 with fc = Identity (src/embeddings.py:112-117); state_dict keys are torchvision's."""

    LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}

    def __init__(self, name):
        super().__init__()
        self.name = name
        self.conv1, self.bn1 = ConvP(3, 64, 7), BNP(64)
        c_in = 64
        for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), self.LAYERS[name])):
            layer = [BasicP(c_in, planes, li > 0)] + [BasicP(planes, planes, False) for _ in range(blocks - 1)]
            setattr(self, f"layer{li + 1}", nn.Sequential(*layer))
            c_in = planes
        self.out_size = 512
