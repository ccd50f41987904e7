"""CLIP RN50 image encoder — drop-in for what the reference gets from `clip.load("RN50")` (src/embeddings.py:305-306)
and calls through `encode_image` (src/embeddings.py:375-376): openai/CLIP's ModifiedResNet (clip/model.py) — a 3-conv
stem, Bottlenecks whose stride is an average pool, and an attention pool (one query = the mean token, 32 heads over the
7 x 7 + 1 tokens) projecting 2048 -> 1024.

The parameter container uses openai/CLIP's key names (`visual.conv1.weight`, `visual.layer3.0.downsample.0.weight`,
`visual.attnpool.positional_embedding`, `visual.attnpool.c_proj.bias`, ...) so CLIP checkpoints interchange; the text
tower is not on the path and is not instantiated. The forward runs in libpvr_b200: the trunk as one encoder program
(program.add_clip_resnet: tcgen05 implicit GEMMs, PVR_OP_AVGPOOL2), then pvr_attnpool_tokens -> one QKV GEMM ->
pvr_attention_mma (50 tokens, head_dim 64) -> the c_proj GEMM on the class-token rows.
"""
import ctypes
import os

import torch
from torch import nn

from .. import _lib
from .. import program as prg
from ..models import gemm
from .moco import _ALLOW_RANDOM_INIT
from .resnet_params import BNP

LAYERS, WIDTH, OUTPUT_DIM, RESOLUTION = (3, 4, 6, 3), 64, 1024, 224  # clip.load("RN50")


class _Conv(nn.Module):
    def __init__(self, c_in, c_out, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c_out, c_in, k, k))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)  # nn.Conv2d default


class _Bottleneck(nn.Module):
    def __init__(self, c_in, planes, stride):
        super().__init__()
        self.conv1, self.bn1 = _Conv(c_in, planes, 1), BNP(planes)
        self.conv2, self.bn2 = _Conv(planes, planes, 3), BNP(planes)
        self.conv3, self.bn3 = _Conv(planes, planes * 4, 1), BNP(planes * 4)
        if stride > 1 or c_in != planes * 4:  # Sequential(OrderedDict("-1": AvgPool2d, "0": conv, "1": bn))
            self.downsample = nn.Sequential()
            self.downsample.add_module("0", _Conv(c_in, planes * 4, 1))
            self.downsample.add_module("1", BNP(planes * 4))


class _AttentionPool(nn.Module):
    def __init__(self, spatial, embed_dim, heads, output_dim):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spatial ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj, self.q_proj, self.v_proj = (nn.Linear(embed_dim, embed_dim) for _ in range(3))
        self.c_proj = nn.Linear(embed_dim, output_dim)
        self.num_heads = heads


class ModifiedResNetParams(nn.Module):
    def __init__(self, layers=LAYERS, output_dim=OUTPUT_DIM, heads=WIDTH * 32 // 64, input_resolution=RESOLUTION,
                 width=WIDTH):
        super().__init__()
        self.output_dim, self.input_resolution, self.layers, self.heads = output_dim, input_resolution, layers, heads
        self.conv1, self.bn1 = _Conv(3, width // 2, 3), BNP(width // 2)
        self.conv2, self.bn2 = _Conv(width // 2, width // 2, 3), BNP(width // 2)
        self.conv3, self.bn3 = _Conv(width // 2, width, 3), BNP(width)
        c_in = width
        for i, blocks in enumerate(layers):
            planes = width * 2 ** i
            seq = [_Bottleneck(c_in, planes, 2 if i > 0 else 1)] + [_Bottleneck(planes * 4, planes, 1)
                                                                      for _ in range(blocks - 1)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*seq))
            c_in = planes * 4
        self.attnpool = _AttentionPool(input_resolution // 32, width * 32, heads, output_dim)
        # CLIP.initialize_parameters: attention-pool projections N(0, embed_dim^-0.5), last BN weight of each block zero
        std = (width * 32) ** -0.5
        for lin in (self.attnpool.q_proj, self.attnpool.k_proj, self.attnpool.v_proj, self.attnpool.c_proj):
            nn.init.normal_(lin.weight, std=std)
        for i in range(4):
            for blk in getattr(self, f"layer{i + 1}"):
                nn.init.zeros_(blk.bn3.weight)


class CLIPResNetModel(nn.Module):
    """`clip.load("RN50")[0]` as far as the reference uses it: `.visual.input_resolution`, `.encode_image(x)`,
    `.parameters()`, `.eval()`, `.to()`."""

    def __init__(self):
        super().__init__()
        self.name = "RN50"
        self.visual = ModifiedResNetParams()
        self.out_size = self.visual.output_dim
        self._runner = None

    def invalidate(self):
        self._runner = None

    def runner(self, device, precision='bf16'):
        if self._runner is None or self._runner.precision != precision or self._runner.device != torch.device(device):
            self._runner = CLIPRNRunner(self.visual, device, precision)
        return self._runner


def load(name, device="cpu", checkpoint_path=None):
    """Stand-in for `clip.load("RN50", device)`: (model, None); see clip_vit.load for the checkpoint handling."""
    if name != "RN50":
        raise NotImplementedError("Requested model not available.")
    model = CLIPResNetModel()
    path = checkpoint_path or "RN50.pt"
    if os.path.isfile(path):
        try:  # openai's published file is a TorchScript archive
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(path, map_location="cpu")
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items() if k.startswith("visual.")}
        model.load_state_dict(sd, strict=True)
    elif not _ALLOW_RANDOM_INIT[-1]:
        raise FileNotFoundError(f"CLIP checkpoint {path} not found (no network access to download it)")
    return model.to(device), None


class CLIPRNRunner:
    """Device state + launch sequence of the RN50 image encoder; precision 'bf16' (tensor cores) or 'fp32' (the parity
    mode: float32 kernels of conv_f32.cu / vit_f32.cu, frames as PVR_FMT_NHWC4_F32)."""

    def __init__(self, vis, device, precision='bf16'):
        self.device, self.precision = torch.device(device), precision
        self.f32 = precision == 'fp32'
        self.lib = _lib.lib()
        self.input_format = _lib.PVR_FMT_NHWC4_F32 if self.f32 else _lib.PVR_FMT_NHWC4_BF16
        self.res, self.heads, self.O = vis.input_resolution, vis.heads, vis.output_dim
        sd = {k: v.detach().cpu() for k, v in vis.state_dict().items()}
        prog = prg.Program()
        hw = self.res
        if self.f32:
            in_slot = prog.new_slot(hw * hw * 4 * 2)
            self.feat_slot, chw = prg.add_clip_resnet_f32(prog, sd, in_slot, hw, vis.layers)
        else:
            in_slot = prog.new_slot(hw * hw * 4)
            self.feat_slot, chw = prg.add_clip_resnet(prog, sd, in_slot, hw, vis.layers)
        prog.emb_width = 1
        self.n_ops = len(prog.ops)
        self.enc = prog.finish(self.device)
        self.C, self.hw = chw[0], chw[1] * chw[2]
        self.S = self.hw + 1
        ap = vis.attnpool
        wdt = torch.float32 if self.f32 else torch.bfloat16
        dev = self.device
        self.pos = ap.positional_embedding.detach().to(dev, torch.float32).contiguous()
        self.wqkv = torch.cat([ap.q_proj.weight, ap.k_proj.weight, ap.v_proj.weight]).detach().to(dev, wdt).contiguous()
        self.bqkv = torch.cat([ap.q_proj.bias, ap.k_proj.bias, ap.v_proj.bias]).detach().to(dev, torch.float32)
        self.wc = ap.c_proj.weight.detach().to(dev, wdt).contiguous()
        self.bc = ap.c_proj.bias.detach().to(dev, torch.float32).contiguous()
        self.n = 0
        C, S = self.C, self.S
        self.flops_per_image = sum(op.get("flops_per_image", 0) for op in prog.ops) + \
            2 * S * C * 3 * C + 4 * S * S * C + 2 * C * self.O

    def bind(self, n):
        if n == self.n:
            return
        dev, dt = self.device, (torch.float32 if self.f32 else torch.bfloat16)
        self.enc.bind(n)
        self.tok = torch.empty(n * self.S, self.C, dtype=dt, device=dev)
        self.qkv = torch.empty(n * self.S, 3 * self.C, dtype=dt, device=dev)
        self.att = torch.empty(n * self.S, self.C, dtype=dt, device=dev)
        self.dummy = torch.zeros(n, 1, device=dev)
        self.n = n

    @property
    def slot0(self):
        return self.enc.slot0

    def launches_per_forward(self):
        return self.n_ops + 4  # trunk ops (upper bound: back-to-back fused pairs are one launch) + attention pool

    def forward(self, out, out_ld=None):
        """Frames must already be in slot0. Writes (n, 1024) fp32 rows into `out`."""
        lib, n, S, C = self.lib, self.n, self.S, self.C
        st = _lib.current_stream_ptr
        ld = out_ld if out_ld is not None else out.stride(0)
        with torch.cuda.device(self.device):
            self.enc.forward(self.dummy, 1)
            feat = self.enc.slot_ptr(self.feat_slot)
            _lib.check(lib.pvr_attnpool_tokens(feat, n, self.hw, C, self.pos.data_ptr(), int(self.f32),
                                               self.tok.data_ptr(), st()), "pvr_attnpool_tokens")
            cls_rows = self.att.view(n, S * C)[:, :C]  # the query token's attention output, row stride S * C
            if self.f32:
                _lib.check(lib.pvr_gemm_f32(self.tok.data_ptr(), C, self.wqkv.data_ptr(), self.bqkv.data_ptr(), None, 0,
                                            self.qkv.data_ptr(), 3 * C, n * S, 3 * C, C, 0, st()), "pvr_gemm_f32")
                _lib.check(lib.pvr_attention_f32(self.qkv.data_ptr(), n, S, C, self.heads, self.att.data_ptr(), st()),
                           "pvr_attention_f32")
                _lib.check(lib.pvr_gemm_f32(cls_rows.data_ptr(), S * C, self.wc.data_ptr(), self.bc.data_ptr(), None, 0,
                                            out.data_ptr(), ld, n, self.O, C, 0, st()), "pvr_gemm_f32")
                return
            gemm(self.tok, self.wqkv, self.qkv, n * S, 3 * C, C, bias=self.bqkv)
            _lib.check(lib.pvr_attention_mma(self.qkv.data_ptr(), n, S, C, self.heads, self.att.data_ptr(), st()),
                       "pvr_attention_mma")
            d = _lib.pvr_gemm_desc()
            d.a, d.lda, d.b, d.ldb = cls_rows.data_ptr(), S * C, self.wc.data_ptr(), C
            d.out, d.ldo, d.bias = out.data_ptr(), ld, self.bc.data_ptr()
            d.m, d.n, d.n_pad, d.k, d.out_f32, d.split_k = n, self.O, self.O, C, 1, 1
            _lib.check(lib.pvr_gemm(ctypes.byref(d), st()), "pvr_gemm")
