"""ORACLE (test infrastructure only — nothing under pvr_habitat_b200/ imports this): CPU restatement of
`EmbeddingNet('maskrcnn_l3')`: src/embeddings.py:283-295 (transforms), :380-383 (`backbone(x)['res4']`),
src/vision_models/maskrcnn.py:26-127 (model construction).

The arithmetic lives in detectron2 (`requirements.txt:18`, `git+https://github.com/facebookresearch/detectron2.git`,
unpinned HEAD), which is NOT installed in this image and not vendored under /root/reference: PARITY UNPINNED against
detectron2's own code. What is restated below, from detectron2's published `modeling/backbone/resnet.py`,
`layers/batch_norm.py`, `layers/wrappers.py` and `config/lazy.py|instantiate.py`:
  * `Conv2d` wrapper = F.conv2d followed by the `norm` sub-module; `FrozenBatchNorm2d` = F.batch_norm on four buffers
    (weight, bias, running_mean, running_var), eps 1e-5, training=False; get_norm("BN") = nn.BatchNorm2d;
  * `BasicStem`: 7x7/2 conv (pad 3, no bias) + norm, ReLU, max_pool2d(3, 2, 1);
  * `BottleneckBlock`: 1x1 -> ReLU -> 3x3 -> ReLU -> 1x1, + shortcut (1x1 conv + norm when in != out), ReLU; with
    `stride_in_1x1=True` the block's stride sits on the FIRST 1x1 conv; `BasicBlock`: 3x3 -> ReLU -> 3x3, + shortcut
    (1x1 conv + norm when in != out), ReLU;
  * `ResNet.make_stage` / `make_default_stages(50)` = [3, 4, 6, 3] blocks, strides [1, 2, 2, 2] on the first block,
    bottleneck_channels = out / 4; `ResNet(stem, stages, out_features=["res4"])` keeps stages res2..res4 only and
    returns {"res4": x};
  * `LazyCall` / `instantiate`: `L(f)(**kw)` records a call, `instantiate` builds the arguments depth first.
`install_detectron2()` registers these under detectron2's module names so that the reference's OWN
`mask_rcnn_model` (stage surgery `stages[-2].extend(...)`, `model.res4[7] = nn.Sequential()`) and transforms run
unmodified on top of them when the goldens are generated (oracle/make_golden.py, build container only). The RPN / ROI
heads never execute on this path and are inert objects.
"""
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from oracle import restate

PIXEL_MEAN = [103.530, 116.280, 123.675]


# ---------------------------------------------------------------------------- detectron2, restated (see header)
class FrozenBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, training=False, eps=self.eps)


def get_norm(norm, out_channels):
    return {"BN": nn.BatchNorm2d, "FrozenBN": FrozenBatchNorm2d}[norm](out_channels)


class Conv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        super().__init__(*args, **kwargs)
        self.norm = norm

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return self.norm(x) if self.norm is not None else x


class BasicStem(nn.Module):
    def __init__(self, in_channels=3, out_channels=64, norm="BN"):
        super().__init__()
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=7, stride=2, padding=3, bias=False,
                            norm=get_norm(norm, out_channels))

    def forward(self, x):
        return F.max_pool2d(F.relu_(self.conv1(x)), kernel_size=3, stride=2, padding=1)


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1, norm="BN",
                 stride_in_1x1=False, dilation=1):
        super().__init__()
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=get_norm(norm, out_channels))
        stride_1x1, stride_3x3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=stride_1x1, bias=False,
                            norm=get_norm(norm, bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=stride_3x3,
                            padding=1 * dilation, bias=False, groups=num_groups, dilation=dilation,
                            norm=get_norm(norm, bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False,
                            norm=get_norm(norm, out_channels))

    def forward(self, x):
        out = self.conv3(F.relu_(self.conv2(F.relu_(self.conv1(x)))))
        out += self.shortcut(x) if self.shortcut is not None else x
        return F.relu_(out)


class BasicBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, stride=1, norm="BN"):
        super().__init__()
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=get_norm(norm, out_channels))
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False,
                            norm=get_norm(norm, out_channels))
        self.conv2 = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False,
                            norm=get_norm(norm, out_channels))

    def forward(self, x):
        out = self.conv2(F.relu_(self.conv1(x)))
        out += self.shortcut(x) if self.shortcut is not None else x
        return F.relu_(out)


class ResNet(nn.Module):
    def __init__(self, stem, stages, num_classes=None, out_features=None, freeze_at=0):
        super().__init__()
        self.stem = stem
        if out_features is not None:  # only the stages that are needed are kept
            stages = stages[:max({"res2": 1, "res3": 2, "res4": 3, "res5": 4}.get(f, 0) for f in out_features)]
        self.stage_names, self.stages = [], []
        for i, blocks in enumerate(stages):
            name = "res" + str(i + 2)
            stage = nn.Sequential(*blocks)
            self.add_module(name, stage)
            self.stage_names.append(name)
            self.stages.append(stage)
        self._out_features = out_features if out_features is not None else [self.stage_names[-1]]

    def forward(self, x):
        assert x.dim() == 4
        outputs = {}
        x = self.stem(x)
        for name, stage in zip(self.stage_names, self.stages):
            x = stage(x)
            if name in self._out_features:
                outputs[name] = x
        return outputs

    @staticmethod
    def make_stage(block_class, num_blocks, *, in_channels, out_channels, **kwargs):
        blocks = []
        for i in range(num_blocks):
            kw = {(k[:-len("_per_block")] if k.endswith("_per_block") else k): (v[i] if k.endswith("_per_block") else v)
                  for k, v in kwargs.items()}
            blocks.append(block_class(in_channels=in_channels, out_channels=out_channels, **kw))
            in_channels = out_channels
        return blocks

    @staticmethod
    def make_default_stages(depth, block_class=None, **kwargs):
        assert depth == 50
        ret = []
        for n, s, i, o in zip([3, 4, 6, 3], [1, 2, 2, 2], [64, 256, 512, 1024], [256, 512, 1024, 2048]):
            kwargs["bottleneck_channels"] = o // 4
            ret.append(ResNet.make_stage(block_class=block_class or BottleneckBlock, num_blocks=n,
                                         stride_per_block=[s] + [1] * (n - 1), in_channels=i, out_channels=o, **kwargs))
        return ret


class GeneralizedRCNN(nn.Module):
    def __init__(self, *, backbone, proposal_generator=None, roi_heads=None, pixel_mean=None, pixel_std=None,
                 input_format=None, vis_period=0):
        super().__init__()
        self.backbone = backbone  # the heads hold no tensors here: they never run on the embedding path


class _LazyCall:
    def __init__(self, target):
        self.target = target

    def __call__(self, **kwargs):
        return {"_target_": self.target, **kwargs}


def instantiate(cfg):
    if isinstance(cfg, dict) and "_target_" in cfg:
        return cfg["_target_"](**{k: instantiate(v) for k, v in cfg.items() if k != "_target_"})
    if isinstance(cfg, (list, tuple)):
        return type(cfg)(instantiate(v) for v in cfg)
    return cfg


class _Inert:
    def __init__(self, *a, **k):
        pass


def install_detectron2():
    """Register the restatement under detectron2's module names (replacing inert stubs of refshim if present)."""
    def inert(attr):  # everything else in the package: an inert class
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (_Inert,), {})

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__getattr__ = inert
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("detectron2")
    mod("detectron2.layers")
    mod("detectron2.config", LazyCall=_LazyCall, instantiate=instantiate)
    mod("detectron2.modeling")
    mod("detectron2.modeling.meta_arch", GeneralizedRCNN=GeneralizedRCNN)
    mod("detectron2.modeling.anchor_generator")
    mod("detectron2.modeling.backbone", BasicStem=BasicStem, BottleneckBlock=BottleneckBlock, ResNet=ResNet)
    mod("detectron2.modeling.backbone.resnet", BasicBlock=BasicBlock)
    for name in ("box_regression", "matcher", "poolers", "proposal_generator", "roi_heads"):
        mod("detectron2.modeling." + name)


# ---------------------------------------------------------------------------- the path, restated
def build_backbone():
    """maskrcnn.py:26-57 + :124-127: R50-C4 stages with res4 extended by BasicBlock(1024, 11), BasicBlock(11, 1024); the
    latter replaced by an empty Sequential."""
    stages = ResNet.make_default_stages(depth=50, stride_in_1x1=True, norm="FrozenBN")
    stages[-2].extend([BasicBlock(1024, 11), BasicBlock(11, 1024)])
    net = ResNet(BasicStem(3, 64, norm="FrozenBN"), stages, out_features=["res4"])
    net.res4[7] = nn.Sequential()
    return net.eval()


def maskrcnn_transforms(frames_nhwc_u8):
    """src/embeddings.py:285-294 after the NHWC -> NCHW transpose of :392: `x[:,:,[0,1,2]] = x[:,:,[2,1,0]]` acts on
    dimension 2 of the NCHW tensor — image ROWS 0 and 2 trade places, the channels stay RGB — then `.float()` (0..255),
    Resize(256) of the FLOAT image (bilinear, no rounding; no antialias under the pinned torchvision 0.10),
    CenterCrop(224), Normalize(pixel mean, 1)."""
    x = np.ascontiguousarray(np.transpose(np.asarray(frames_nhwc_u8), (0, 3, 1, 2)))
    x[:, :, [0, 1, 2]] = x[:, :, [2, 1, 0]]
    h, w = x.shape[2:]
    rh, rw, top, left = restate.resize_geometry(h, w, 256, 224)
    y = restate.resize_bilinear_f32(x, rh, rw)[:, :, top:top + 224, left:left + 224]
    mean = np.asarray(PIXEL_MEAN, np.float32)[None, :, None, None]
    return ((y - mean) / np.float32(1.0)).astype(np.float32)


def embedding_forward(sd, frames_nhwc_u8):
    """(N, H, W, 3) uint8 -> (N, 2156) float32; `sd`: the backbone's state_dict (detectron2 key names)."""
    net = build_backbone()
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = net(torch.from_numpy(maskrcnn_transforms(frames_nhwc_u8)))["res4"]
    return out.reshape(out.shape[0], -1).numpy()


def maskrcnn_state(seed):
    """Deterministic backbone weights under detectron2's key names (numpy default_rng: platform independent): kaiming
    fan-out convolutions and non-trivial norm affines / statistics like restate.resnet50_state (gain 0.5 on a block's
    last norm, 0.7 on the shortcut's). Activations stay O(1) on the 0..255 input scale because the stem's normalisation
    statistics are set for that scale (running_var ~ the variance of a 7x7 response to pixels of spread ~60)."""
    rng = np.random.default_rng(seed)
    ref = build_backbone().state_dict()
    sd = {}
    for k, v in ref.items():
        gain = 0.5 if ".conv3.norm." in k else (0.7 if ".shortcut.norm." in k else 1.0)
        if k.endswith("num_batches_tracked"):
            sd[k] = v.clone()
        elif k.endswith("running_var"):
            scale = 100.0 if k.startswith("stem.") else 1.0
            sd[k] = torch.from_numpy((scale * rng.uniform(0.75, 1.25, tuple(v.shape))).astype(np.float32))
        elif k.endswith("running_mean"):
            sd[k] = torch.from_numpy((0.1 * rng.standard_normal(tuple(v.shape))).astype(np.float32))
        elif k.endswith("norm.weight"):
            sd[k] = torch.from_numpy(rng.uniform(0.75 * gain, 1.25 * gain, tuple(v.shape)).astype(np.float32))
        elif k.endswith("norm.bias"):
            sd[k] = torch.from_numpy((0.1 * rng.standard_normal(tuple(v.shape))).astype(np.float32))
        else:
            fan_out = v.shape[0] * v.shape[2] * v.shape[3]
            sd[k] = torch.from_numpy((rng.standard_normal(tuple(v.shape)) * (2.0 / fan_out) ** 0.5).astype(np.float32))
    return sd
