"""ORACLE — test infrastructure only (nothing under pvr_habitat_b200/ imports this). CPU restatement (torch fp32) of the
CLIP RN50 image encoder the reference reaches through `clip.load("RN50")[0].encode_image` (src/embeddings.py:305-306,
375-376) behind CLIP's transforms (src/embeddings.py:309-314).

The arithmetic lives in openai/CLIP (`clip/model.py`: Bottleneck, AttentionPool2d, ModifiedResNet; unpinned HEAD in
requirements.txt:19), which is neither vendored in the reference nor installed here, and `transformers` has no
CLIP-ResNet to cross-check against: PARITY UNPINNED against openai/CLIP itself. Published algorithm restated below as
torch modules with openai/CLIP's parameter names:
  * stem: conv 3x3/2 (3 -> 32) + BN + ReLU, conv 3x3 (32 -> 32) + BN + ReLU, conv 3x3 (32 -> 64) + BN + ReLU, AvgPool2d(2);
  * Bottleneck(inplanes, planes, stride): conv1 1x1 + BN + ReLU, conv2 3x3 (stride 1) + BN + ReLU, AvgPool2d(stride) when
    stride > 1, conv3 1x1 + BN; shortcut = AvgPool2d(stride) -> conv 1x1 (stride 1) -> BN when stride > 1 or
    inplanes != 4 planes; ReLU(out + shortcut); layers (3, 4, 6, 3), widths 64 / 128 / 256 / 512 (x 4), strides 1 / 2 / 2 / 2;
  * AttentionPool2d(7, 2048, 32 heads, 1024): tokens = [mean over positions | positions] + positional_embedding;
    F.multi_head_attention_forward with query = the mean token, separate q / k / v projections (with biases),
    out_proj = c_proj; returns the single query row.
`FakeClip` wraps the model the way `clip.load` returns it (`.visual.input_resolution`, `.encode_image`) so that the
reference's own EmbeddingNet('clip_rn50') — transforms, `encode_image` dispatch, output reshaping — runs unmodified on
top of it when the goldens are generated (oracle/make_golden.py, build container only).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)),
                ("0", nn.Conv2d(inplanes, planes * 4, 1, stride=1, bias=False)),
                ("1", nn.BatchNorm2d(planes * 4))]))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(self.avgpool(out)))
        identity = self.downsample(x) if self.downsample is not None else x
        return F.relu(out + identity)


class AttentionPool2d(nn.Module):
    def __init__(self, spacial_dim, embed_dim, num_heads, output_dim=None):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads

    def forward(self, x):
        x = x.flatten(start_dim=2).permute(2, 0, 1)  # NCHW -> (HW)NC
        x = torch.cat([x.mean(dim=0, keepdim=True), x], dim=0)  # (HW+1)NC
        x = x + self.positional_embedding[:, None, :].to(x.dtype)
        x, _ = F.multi_head_attention_forward(
            query=x[:1], key=x, value=x, embed_dim_to_check=x.shape[-1], num_heads=self.num_heads,
            q_proj_weight=self.q_proj.weight, k_proj_weight=self.k_proj.weight, v_proj_weight=self.v_proj.weight,
            in_proj_weight=None, in_proj_bias=torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias]),
            bias_k=None, bias_v=None, add_zero_attn=False, dropout_p=0, out_proj_weight=self.c_proj.weight,
            out_proj_bias=self.c_proj.bias, use_separate_proj_weight=True, training=self.training, need_weights=False)
        return x.squeeze(0)


class ModifiedResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), output_dim=1024, heads=32, input_resolution=224, width=64):
        super().__init__()
        self.output_dim, self.input_resolution = output_dim, input_resolution
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        self.avgpool = nn.AvgPool2d(2)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        self.attnpool = AttentionPool2d(input_resolution // 32, width * 32, heads, output_dim)

    def _make_layer(self, planes, blocks, stride=1):
        layers = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        layers += [Bottleneck(self._inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = x.type(self.conv1.weight.dtype)
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.relu(bn(conv(x)))
        x = self.avgpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.attnpool(x)


class FakeClip(nn.Module):
    """What `clip.load("RN50", device)` returns, as far as src/embeddings.py uses it."""

    def __init__(self):
        super().__init__()
        self.visual = ModifiedResNet()

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))


def clip_rn50_state(seed):
    """Deterministic weights (numpy default_rng) under openai/CLIP's `visual.*` key names: kaiming fan-out convolutions,
    non-trivial BN affines / statistics (gain 0.5 on a block's last BN, 0.7 on the shortcut's — CLIP's own
    initialisation zeroes bn3.weight, which would switch the residual branches off), attention-pool projections
    N(0, 2048^-0.5) with small biases."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in FakeClip().state_dict().items():
        shape = tuple(v.shape)
        gain = 0.5 if ".bn3." in k and "layer" in k else (0.7 if ".downsample.1." in k else 1.0)
        if k.endswith("num_batches_tracked"):
            t = v.clone()
        elif k.endswith("running_var"):
            t = rng.uniform(0.75, 1.25, shape)
        elif k.endswith("running_mean") or (k.endswith(".bias") and "bn" in k.split(".")[-2] or ".downsample.1.bias" in k):
            t = 0.1 * rng.standard_normal(shape)
        elif k.endswith(".weight") and v.dim() == 1:
            t = rng.uniform(0.75 * gain, 1.25 * gain, shape)
        elif v.dim() == 4:
            t = rng.standard_normal(shape) * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
        elif k.endswith("positional_embedding"):
            t = rng.standard_normal(shape) * 0.5
        elif v.dim() == 2:
            t = rng.standard_normal(shape) * shape[1] ** -0.5
        else:  # Linear biases
            t = 0.02 * rng.standard_normal(shape)
        sd[k] = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.asarray(t, dtype=np.float32))
    return sd


def embedding_forward(sd, frames_nhwc_u8):
    """(N, H, W, 3) uint8 -> (N, 1024) float32; `sd` with `visual.*` keys."""
    from oracle import restate_vit
    m = FakeClip().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        return m.encode_image(torch.from_numpy(restate_vit.clip_transforms(frames_nhwc_u8))).numpy()
