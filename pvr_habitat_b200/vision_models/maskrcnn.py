"""Mask R-CNN R50-C4 backbone with the 1024 -> 11 compression block — drop-in for what the reference builds in
src/vision_models/maskrcnn.py:26-127 (`mask_rcnn_model`) and keeps as `EmbeddingNet('maskrcnn_l3').embedding`:
`GeneralizedRCNN(...).backbone` with `res4[7] = nn.Sequential()`, evaluated as `backbone(x)['res4']`
(src/embeddings.py:380-383) -> (N, 11, 14, 14) -> 2156 values per frame.

detectron2 is not a dependency here: the module holds the parameters under detectron2's `ResNet` key names
(`stem.conv1.weight`, `stem.conv1.norm.running_var`, `res2.0.shortcut.norm.bias`, `res4.6.conv2.weight`, ...) so that a
reference checkpoint's `model` dict loads (everything outside `backbone.` — RPN, ROI heads — and the dropped `res4.7`
block is ignored), and translates them into the torchvision naming that program.add_resnet50 reads. The arithmetic is
the same CUDA program as the MoCo layer-3 encoders with two differences: the stride of a down-sampling bottleneck sits
on its first 1x1 convolution (`stride_in_1x1=True`, maskrcnn.py:52-56), and the compression block's shortcut is an
unbiased 1x1 convolution (detectron2 BasicBlock) instead of a biased 3x3 one — expressed as a 3x3 kernel with only its
centre tap set, which is exact.
"""
import os

import torch
from torch import nn

from .moco import _ALLOW_RANDOM_INIT
from .resnet_params import BNP, ConvP

PIXEL_MEAN = [103.530, 116.280, 123.675]  # maskrcnn.py:120 / src/embeddings.py:293 (applied to R, G, B: see Transforms)
_STAGES = (("res2", 64, 3), ("res3", 128, 4), ("res4", 256, 6))


class _FrozenBN(nn.Module):
    """detectron2 FrozenBatchNorm2d: four buffers, eps 1e-5, no num_batches_tracked."""

    def __init__(self, c):
        super().__init__()
        self.register_buffer("weight", torch.ones(c))
        self.register_buffer("bias", torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c) - 1e-5)


class _Conv(ConvP):
    """detectron2 Conv2d wrapper: the normalisation layer is the sub-module `norm` of the convolution."""

    def __init__(self, c_in, c_out, k, frozen=True):
        super().__init__(c_in, c_out, k)
        self.norm = _FrozenBN(c_out) if frozen else BNP(c_out)


class _Bottleneck(nn.Module):
    def __init__(self, c_in, planes, shortcut):
        super().__init__()
        if shortcut:
            self.shortcut = _Conv(c_in, planes * 4, 1)
        self.conv1, self.conv2, self.conv3 = _Conv(c_in, planes, 1), _Conv(planes, planes, 3), _Conv(planes, planes * 4, 1)


class _Basic(nn.Module):
    """detectron2 BasicBlock(1024, 11) of make_compress_stages (maskrcnn.py:26-45): norm="BN" (the make_stage default)."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.shortcut = _Conv(c_in, c_out, 1, frozen=False)
        self.conv1, self.conv2 = _Conv(c_in, c_out, 3, frozen=False), _Conv(c_out, c_out, 3, frozen=False)


class _Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = _Conv(3, 64, 7)


class MaskRCNNBackboneParams(nn.Module):
    variant = "l3"
    stride_in_1x1 = True

    def __init__(self):
        super().__init__()
        self.stem = _Stem()
        c_in = 64
        for name, planes, blocks in _STAGES:
            layer = [_Bottleneck(c_in, planes, True)] + [_Bottleneck(planes * 4, planes, False) for _ in range(blocks - 1)]
            if name == "res4":
                layer += [_Basic(1024, 11), nn.Sequential()]  # res4[6] compresses; res4[7] is emptied by the reference
            setattr(self, name, nn.Sequential(*layer))
            c_in = planes * 4
        self.out_size = 11 * 14 * 14

    def forward(self, *a, **k):
        from .. import _lib
        raise _lib.PvrError("MaskRCNNBackboneParams holds parameters only; use EmbeddingNet (CUDA program)")

    @staticmethod
    def program_state_dict(sd):
        """detectron2 names -> the torchvision names of program.add_resnet50(variant='l3')."""
        out = {}

        def conv(src, dst_conv, dst_bn):
            out[dst_conv + ".weight"] = sd[src + ".weight"]
            for k in ("weight", "bias", "running_mean", "running_var"):
                out[f"{dst_bn}.{k}"] = sd[f"{src}.norm.{k}"]

        conv("stem.conv1", "conv1", "bn1")
        for (name, _, blocks), layer in zip(_STAGES, ("layer1.", "layer2.", "layer3.0.")):
            for b in range(blocks):
                for i in (1, 2, 3):
                    conv(f"{name}.{b}.conv{i}", f"{layer}{b}.conv{i}", f"{layer}{b}.bn{i}")
                if b == 0:
                    conv(f"{name}.0.shortcut", f"{layer}0.downsample.0", f"{layer}0.downsample.1")
        conv("res4.6.conv1", "layer3.1.conv1", "layer3.1.bn1")
        conv("res4.6.conv2", "layer3.1.conv2", "layer3.1.bn2")
        conv("res4.6.shortcut", "layer3.1.downsample.0", "layer3.1.downsample.1")
        w1 = sd["res4.6.shortcut.weight"]  # (11, 1024, 1, 1) -> centre tap of a 3x3 kernel, zero bias
        w3 = torch.zeros(w1.shape[0], w1.shape[1], 3, 3, dtype=w1.dtype)
        w3[:, :, 1, 1] = w1[:, :, 0, 0]
        out["layer3.1.downsample.0.weight"] = w3
        out["layer3.1.downsample.0.bias"] = torch.zeros(w1.shape[0])
        return out


def mask_rcnn_model(checkpoint_path):
    """`mask_rcnn_model(checkpoint_path)` of maskrcnn.py:59-127: `torch.load(path)['model']` is the state_dict of the
    whole GeneralizedRCNN (loaded strictly there); the backbone's entries are taken, `res4.7` and the heads dropped."""
    model = MaskRCNNBackboneParams()
    if os.path.isfile(checkpoint_path):
        full = torch.load(checkpoint_path, map_location="cpu")["model"]
        sd = {k[len("backbone."):]: v for k, v in full.items()
              if k.startswith("backbone.") and not k.startswith("backbone.res4.7.")}
        model.load_state_dict(sd, strict=True)
    elif not _ALLOW_RANDOM_INIT[-1]:
        raise FileNotFoundError(f"No such file or directory: '{checkpoint_path}'")
    return model
