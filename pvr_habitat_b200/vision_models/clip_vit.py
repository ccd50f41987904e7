"""CLIP-architecture ViT-B image encoder — drop-in for what the reference gets from `clip.load("ViT-B/32")`
(src/embeddings.py:298-314) and calls through `encode_image` (src/embeddings.py:375-376).

The parameter container uses openai/CLIP's key names (`visual.conv1.weight`, `visual.class_embedding`,
`visual.transformer.resblocks.N.attn.in_proj_weight`, ..., `visual.proj`) so CLIP checkpoints interchange; the text
tower is not on the path and is not instantiated. The forward runs in libpvr_b200: patch embedding as an implicit GEMM
straight from the NHWC4 frames (im2col TMA: one "pixel" = one patch row), fused token assembly + ln_pre, LayerNorm,
tcgen05 attention, tcgen05 GEMMs with bias / QuickGELU / fp32-residual epilogues.
"""
import ctypes
import os

import torch
from torch import nn

from .. import _lib
from .. import program as prg
from ..models import gemm
from .moco import _ALLOW_RANDOM_INIT

_CONFIGS = {"ViT-B/32": 32, "ViT-B/16": 16}


class _LN(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(w))
        self.bias = nn.Parameter(torch.zeros(w))


class _Lin(nn.Module):
    def __init__(self, i, o, std):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(o, i) * std)
        self.bias = nn.Parameter(torch.zeros(o))


class _Attn(nn.Module):
    def __init__(self, w, attn_std, proj_std):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.randn(3 * w, w) * attn_std)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * w))
        self.out_proj = _Lin(w, w, proj_std)


class _MLP(nn.Module):
    def __init__(self, w, fc_std, proj_std):
        super().__init__()
        self.c_fc = _Lin(w, 4 * w, fc_std)
        self.c_proj = _Lin(4 * w, w, proj_std)


class _Block(nn.Module):
    def __init__(self, w, attn_std, proj_std, fc_std):
        super().__init__()
        self.attn = _Attn(w, attn_std, proj_std)
        self.ln_1 = _LN(w)
        self.mlp = _MLP(w, fc_std, proj_std)
        self.ln_2 = _LN(w)


class _Transformer(nn.Module):
    def __init__(self, w, layers):
        super().__init__()
        proj_std = (w ** -0.5) * ((2 * layers) ** -0.5)  # CLIP.initialize_parameters
        self.resblocks = nn.Sequential(*[_Block(w, w ** -0.5, proj_std, (2 * w) ** -0.5) for _ in range(layers)])


class _Conv1(nn.Module):
    def __init__(self, w, patch):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(w, 3, patch, patch))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)


class VisionTransformerParams(nn.Module):
    def __init__(self, input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        assert width == 768 and heads * 64 == width, "libpvr_b200 ViT kernels are built for ViT-B (width 768)"
        self.input_resolution, self.patch_size, self.width = input_resolution, patch_size, width
        self.layers, self.heads, self.output_dim = layers, heads, output_dim
        scale = width ** -0.5
        self.conv1 = _Conv1(width, patch_size)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = _LN(width)
        self.transformer = _Transformer(width, layers)
        self.ln_post = _LN(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


class CLIPImageModel(nn.Module):
    """`clip.load(...)[0]` as far as the reference uses it: `.visual.input_resolution`, `.encode_image(x)`,
    `.parameters()`, `.eval()`, `.to()`."""

    def __init__(self, name="ViT-B/32"):
        super().__init__()
        self.name = name
        self.visual = VisionTransformerParams(patch_size=_CONFIGS[name])
        self.out_size = self.visual.output_dim
        self._runner = None

    def invalidate(self):
        self._runner = None

    def runner(self, device, precision='bf16'):
        cls = ViTRunnerF32 if precision == 'fp32' else ViTRunner
        if type(self._runner) is not cls or self._runner.device != torch.device(device):
            self._runner = cls(self.visual, device)
        return self._runner


def load(name, device="cpu", checkpoint_path=None):
    """Stand-in for `clip.load(name, device)`: returns (model, None). openai's checkpoints are downloaded TorchScript
    archives; offline a state_dict file may be given, otherwise (inside `allow_random_init()`) the weights stay at
    CLIP's random initialisation."""
    if name not in _CONFIGS:
        raise NotImplementedError("Requested model not available.")
    model = CLIPImageModel(name)
    path = checkpoint_path or (name.replace("/", "-") + ".pt")
    if os.path.isfile(path):
        try:  # openai's published files ("ViT-B-32.pt") are TorchScript archives
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:  # a plain state_dict file
            sd = torch.load(path, map_location="cpu")
        sd = {k: v.float() for k, v in sd.items() if k.startswith("visual.")}
        model.load_state_dict(sd, strict=True)
    elif not _ALLOW_RANDOM_INIT[-1]:
        raise FileNotFoundError(f"CLIP checkpoint {path} not found (no network access to download it)")
    return model.to(device), None


def pack_patch_weight(w, dtype=torch.bfloat16):
    """(width, 3, p, p) -> (width, p * p * 4): K ordered (patch row, pixel, channel padded to 4)."""
    co, ci, p, _ = w.shape
    out = torch.zeros(co, p, p, 4, dtype=torch.float32)
    out[..., :3] = w.permute(0, 2, 3, 1)
    return out.reshape(co, p * p * 4).to(dtype)


def clip_spec(vis):
    """The pieces of a CLIP VisionTransformer as ViTRunner consumes them."""
    blocks = [dict(ln1=(blk.ln_1.weight, blk.ln_1.bias), ln2=(blk.ln_2.weight, blk.ln_2.bias),
                   wqkv=blk.attn.in_proj_weight, bqkv=blk.attn.in_proj_bias,
                   wo=blk.attn.out_proj.weight, bo=blk.attn.out_proj.bias,
                   w1=blk.mlp.c_fc.weight, b1=blk.mlp.c_fc.bias, w2=blk.mlp.c_proj.weight, b2=blk.mlp.c_proj.bias)
              for blk in vis.transformer.resblocks]
    return dict(patch=vis.patch_size, width=vis.width, heads=vis.heads, resolution=vis.input_resolution,
                patch_weight=vis.conv1.weight, patch_bias=None, cls=vis.class_embedding, pos=vis.positional_embedding,
                ln_pre=(vis.ln_pre.weight, vis.ln_pre.bias), blocks=blocks,
                ln_post=(vis.ln_post.weight, vis.ln_post.bias), proj=vis.proj, eps=1e-5, act=2)


class ViTRunner:
    """Device state (bf16 weights, buffers) + the launch sequence of one pre-LN ViT image encoder.

    `spec` (see clip_spec / mae.mae_spec) names the tensors: CLIP has ln_pre, QuickGELU (act 2), eps 1e-5 and a final
    projection of the normalised class token; MAE (timm blocks) has a biased patch embedding, no ln_pre, erf GELU
    (act 3), eps 1e-6 and returns the normalised class token itself in float32."""

    def __init__(self, spec, device):
        if not isinstance(spec, dict):
            spec = clip_spec(spec)
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self.input_format = _lib.PVR_FMT_NHWC4_BF16
        self.p, self.W, self.heads = spec["patch"], spec["width"], spec["heads"]
        self.L = len(spec["blocks"])
        self.res = spec["resolution"]
        self.eps, self.act = float(spec["eps"]), int(spec["act"])
        self.grid = self.res // self.p
        self.S = self.grid * self.grid + 1
        dev, bf = self.device, torch.bfloat16
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
        b = lambda t: t.detach().to(dev, bf).contiguous()  # noqa: E731
        # Patch embedding as a conv op of the encoder program. TMA traversal strides stop at 8, so the stride-p conv
        # is expressed on a reshaped view of the NHWC4 frames: (N*grid, p, grid, p*4) = (patch row, row inside the
        # patch, patch column, one patch row of pixels). One "image" of the program is one row of patches, the filter
        # is p x 1 taps over the "row inside the patch" axis and every stride is 1.
        # Patch sizes whose rows are not whole 128-byte TMA rows (mae_huge: p = 14) gather their patches with
        # pvr_vit_patchify and multiply them by the weight in its torch layout (K = 3 p^2 padded to a multiple of 64).
        prog = prg.Program()
        cpp = self.p * 4
        pbias = spec["patch_bias"]
        pbias = torch.zeros(self.W) if pbias is None else pbias.detach().cpu().float()
        self.gather_patches = (cpp * 2) % 128 != 0
        if self.gather_patches:
            self.patch_enc = None  # the frames live in a plain buffer (bind), no encoder program
            k = 3 * self.p * self.p
            self.patch_k = (k + 63) // 64 * 64
            wp = torch.zeros(self.W, self.patch_k)
            wp[:, :k] = spec["patch_weight"].detach().cpu().float().reshape(self.W, k)
            self.patch_w, self.patch_b = wp.to(dev, bf).contiguous(), pbias.to(dev)
        else:
            self.in_slot = prog.new_slot(self.p * self.grid * cpp)
            self.patch_slot = prog.conv(self.in_slot, (cpp, self.p, self.grid),
                                        pack_patch_weight(spec["patch_weight"].detach().cpu().float()), self.p * cpp,
                                        self.W, self.p, 1, (1, 1), (0, 0), (1, self.grid), torch.ones(self.W), pbias, 0,
                                        flops=2 * self.grid * self.W * 3 * self.p * self.p)
            prog.emb_width = 1
            self.patch_enc = prog.finish(dev)
        self.cls, self.pos = f(spec["cls"].reshape(-1)), f(spec["pos"].reshape(self.S, self.W))
        self.ln_pre = tuple(f(t) for t in spec["ln_pre"]) if spec["ln_pre"] is not None else None
        self.ln_post = tuple(f(t) for t in spec["ln_post"])
        self.blocks = []
        for blk in spec["blocks"]:
            self.blocks.append(dict(
                ln1=tuple(f(t) for t in blk["ln1"]), ln2=tuple(f(t) for t in blk["ln2"]),
                wqkv=b(blk["wqkv"]), bqkv=f(blk["bqkv"]), wo=b(blk["wo"]), bo=f(blk["bo"]),
                w1=b(blk["w1"]), b1=f(blk["b1"]), w2=b(blk["w2"]), b2=f(blk["b2"])))
        self.proj_t = b(spec["proj"].t()) if spec["proj"] is not None else None  # (output_dim, width): K-major B
        self.O = self.proj_t.shape[0] if self.proj_t is not None else self.W
        self.n = 0
        self.flops_per_image = (2 * self.grid ** 2 * self.W * 3 * self.p ** 2
                                + self.L * (2 * self.S * self.W * 12 * self.W + 4 * self.S * self.S * self.W)
                                + (2 * self.W * self.O if self.proj_t is not None else 0))

    def bind(self, n):
        if n == self.n:
            return
        dev, bf, M, W = self.device, torch.bfloat16, n * self.S, self.W
        if self.gather_patches:
            self.frames = torch.empty(n, self.res, self.res, 4, dtype=bf, device=dev)
        else:
            self.patch_enc.bind(n * self.grid)
        self.x = torch.empty(M, W, dtype=torch.float32, device=dev)
        self.y = torch.empty(M, W, dtype=bf, device=dev)
        self.qkv = torch.empty(M, 3 * W, dtype=bf, device=dev)
        self.att = torch.empty(M, W, dtype=bf, device=dev)
        self.h = torch.empty(M, 4 * W, dtype=bf, device=dev)
        self.clsy = torch.empty(n, W, dtype=bf, device=dev)
        self.dummy = torch.zeros(n * self.grid, 1, device=dev)
        if self.gather_patches:
            self.col = torch.empty(n * self.grid ** 2, self.patch_k, dtype=bf, device=dev)
            self.patches = torch.empty(n * self.grid ** 2, W, dtype=bf, device=dev)
        self.n = n

    @property
    def slot0(self):
        return self.frames.data_ptr() if self.gather_patches else self.patch_enc.slot0

    def launches_per_forward(self):
        return (3 if self.gather_patches else 2) + 7 * self.L + (2 if self.proj_t is not None else 1)

    def forward(self, out, out_ld=None):
        """Frames must already be in slot0 (NHWC4 bf16). Writes (n, O) fp32 rows into `out`."""
        lib, n, S, W, M, eps = self.lib, self.n, self.S, self.W, self.n * self.S, self.eps
        st = _lib.current_stream_ptr
        ld = out_ld if out_ld is not None else out.stride(0)
        with torch.cuda.device(self.device):
            if self.gather_patches:
                _lib.check(lib.pvr_vit_patchify(self.frames.data_ptr(), n, self.res, self.p, self.patch_k, 0,
                                                self.col.data_ptr(), st()), "pvr_vit_patchify")
                gemm(self.col, self.patch_w, self.patches, n * self.grid ** 2, W, self.patch_k, bias=self.patch_b)
                patches = self.patches.data_ptr()
            else:
                self.patch_enc.forward(self.dummy, 1)
                patches = self.patch_enc.slot_ptr(self.patch_slot)
            g, bta = (self.ln_pre[0].data_ptr(), self.ln_pre[1].data_ptr()) if self.ln_pre is not None else (None, None)
            _lib.check(lib.pvr_vit_embed(patches, self.cls.data_ptr(), self.pos.data_ptr(), n, S, W, g, bta, eps,
                                         self.x.data_ptr(), st()), "pvr_vit_embed")
            for blk in self.blocks:
                _lib.check(lib.pvr_layernorm(self.x.data_ptr(), 1, M, W, blk["ln1"][0].data_ptr(),
                                             blk["ln1"][1].data_ptr(), eps, self.y.data_ptr(), st()), "pvr_layernorm")
                gemm(self.y, blk["wqkv"], self.qkv, M, 3 * W, W, bias=blk["bqkv"])
                _lib.check(lib.pvr_attention(self.qkv.data_ptr(), n, S, W, self.heads, self.att.data_ptr(), st()),
                           "pvr_attention")
                gemm(self.att, blk["wo"], self.x, M, W, W, bias=blk["bo"], res=self.x, out_f32=1)
                _lib.check(lib.pvr_layernorm(self.x.data_ptr(), 1, M, W, blk["ln2"][0].data_ptr(),
                                             blk["ln2"][1].data_ptr(), eps, self.y.data_ptr(), st()), "pvr_layernorm")
                gemm(self.y, blk["w1"], self.h, M, 4 * W, W, bias=blk["b1"], act=self.act)
                gemm(self.h, blk["w2"], self.x, M, W, 4 * W, bias=blk["b2"], res=self.x, out_f32=1)
            if self.proj_t is None:  # MAE: the normalised class token is the embedding (float32)
                if ld != W:
                    raise _lib.PvrError("ViTRunner: the embedding rows must be dense (ld == width)")
                _lib.check(lib.pvr_layernorm_f32(self.x.data_ptr(), S, n, W, self.ln_post[0].data_ptr(),
                                                 self.ln_post[1].data_ptr(), eps, out.data_ptr(), W, st()),
                           "pvr_layernorm_f32")
                return
            _lib.check(lib.pvr_layernorm(self.x.data_ptr(), S, n, W, self.ln_post[0].data_ptr(),
                                         self.ln_post[1].data_ptr(), eps, self.clsy.data_ptr(), st()), "pvr_layernorm")
            d = _lib.pvr_gemm_desc()
            d.a, d.lda, d.b, d.ldb = self.clsy.data_ptr(), W, self.proj_t.data_ptr(), W
            d.out, d.ldo = out.data_ptr(), ld
            d.m, d.n, d.n_pad, d.k, d.out_f32, d.split_k = n, self.O, self.O, W, 1, 1
            _lib.check(lib.pvr_gemm(ctypes.byref(d), st()), "pvr_gemm")


class ViTRunnerF32:
    """fp32 parity mode of ViTRunner (north star: embeddings within 1e-5 relative L2 "in the fp32 mode"): the same
    launch sequence with float32 weights, activations and accumulation on the CUDA cores (csrc/vit_f32.cu, conv_f32.cu).
    Frames arrive as PVR_FMT_NHWC4_F32. A checking mode, not a performance path."""

    def __init__(self, spec, device):
        if not isinstance(spec, dict):
            spec = clip_spec(spec)
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self.input_format = _lib.PVR_FMT_NHWC4_F32
        self.p, self.W, self.heads = spec["patch"], spec["width"], spec["heads"]
        self.L = len(spec["blocks"])
        self.res = spec["resolution"]
        self.eps, self.act = float(spec["eps"]), int(spec["act"])
        self.grid = self.res // self.p
        self.S = self.grid * self.grid + 1
        dev = self.device
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
        prog = prg.Program()
        cpp = self.p * 4
        self.in_slot = prog.new_slot(self.p * self.grid * cpp * 2)  # float32 values: two bf16 elements each
        pbias = spec["patch_bias"]
        pbias = torch.zeros(self.W) if pbias is None else pbias.detach().cpu().float()
        self.gather_patches = (cpp * 2) % 128 != 0  # same split as ViTRunner (mae_huge: p = 14)
        if self.gather_patches:
            self.patch_enc = None
            k = 3 * self.p * self.p
            self.patch_k = (k + 63) // 64 * 64
            wp = torch.zeros(self.W, self.patch_k)
            wp[:, :k] = spec["patch_weight"].detach().cpu().float().reshape(self.W, k)
            self.patch_w, self.patch_b = wp.to(dev).contiguous(), pbias.to(dev)
        else:
            self.patch_slot = prog.alloc(self.grid * self.W * 2)
            prog.conv(self.in_slot, (cpp, self.p, self.grid),
                      pack_patch_weight(spec["patch_weight"].detach().cpu().float(), torch.float32), self.p * cpp,
                      self.W, self.p, 1, (1, 1), (0, 0), (1, self.grid), torch.ones(self.W), pbias, 0,
                      out_slot=self.patch_slot, out_pitch=self.W, flags=prg.F32)
            prog.emb_width = 1
            self.patch_enc = prog.finish(dev)
        self.cls, self.pos = f(spec["cls"].reshape(-1)), f(spec["pos"].reshape(self.S, self.W))
        self.ln_pre = tuple(f(t) for t in spec["ln_pre"]) if spec["ln_pre"] is not None else None
        self.ln_post = tuple(f(t) for t in spec["ln_post"])
        self.blocks = [{k: (tuple(f(t) for t in v) if isinstance(v, tuple) else f(v)) for k, v in blk.items()}
                       for blk in spec["blocks"]]
        self.proj_t = f(spec["proj"].t()) if spec["proj"] is not None else None
        self.O = self.proj_t.shape[0] if self.proj_t is not None else self.W
        self.zero_bias = torch.zeros(max(self.O, 1), device=dev)
        self.n = 0
        self.flops_per_image = 0

    def bind(self, n):
        if n == self.n:
            return
        dev, M, W = self.device, n * self.S, self.W
        if self.gather_patches:
            self.frames = torch.empty(n, self.res, self.res, 4, dtype=torch.float32, device=dev)
        else:
            self.patch_enc.bind(n * self.grid)
        f32 = torch.float32
        self.x = torch.empty(M, W, dtype=f32, device=dev)
        self.y = torch.empty(M, W, dtype=f32, device=dev)
        self.qkv = torch.empty(M, 3 * W, dtype=f32, device=dev)
        self.att = torch.empty(M, W, dtype=f32, device=dev)
        self.h = torch.empty(M, 4 * W, dtype=f32, device=dev)
        self.clsy = torch.empty(n, W, dtype=f32, device=dev)
        self.dummy = torch.zeros(n * self.grid, 1, device=dev)
        if self.gather_patches:
            self.col = torch.empty(n * self.grid ** 2, self.patch_k, dtype=f32, device=dev)
            self.patches = torch.empty(n * self.grid ** 2, W, dtype=f32, device=dev)
        self.n = n

    @property
    def slot0(self):
        return self.frames.data_ptr() if self.gather_patches else self.patch_enc.slot0

    def launches_per_forward(self):
        return (3 if self.gather_patches else 2) + 7 * self.L + 2

    def _gemm(self, a, w, bias, out, M, N, K, res=None, act=0):
        _lib.check(self.lib.pvr_gemm_f32(a.data_ptr(), a.stride(0), w.data_ptr(), bias.data_ptr(),
                                         res.data_ptr() if res is not None else None,
                                         res.stride(0) if res is not None else 0, out.data_ptr(), out.stride(0), M, N, K,
                                         act, _lib.current_stream_ptr()), "pvr_gemm_f32")

    def _ln(self, x, row_step, rows, gb, out):
        _lib.check(self.lib.pvr_layernorm_f32(x.data_ptr(), row_step, rows, self.W, gb[0].data_ptr(), gb[1].data_ptr(),
                                              self.eps, out.data_ptr(), self.W, _lib.current_stream_ptr()),
                   "pvr_layernorm_f32")

    def forward(self, out, out_ld=None):
        lib, n, S, W, M = self.lib, self.n, self.S, self.W, self.n * self.S
        ld = out_ld if out_ld is not None else out.stride(0)
        with torch.cuda.device(self.device):
            if self.gather_patches:
                _lib.check(lib.pvr_vit_patchify(self.frames.data_ptr(), n, self.res, self.p, self.patch_k, 1,
                                                self.col.data_ptr(), _lib.current_stream_ptr()), "pvr_vit_patchify")
                self._gemm(self.col, self.patch_w, self.patch_b, self.patches, n * self.grid ** 2, W, self.patch_k)
                patches = self.patches.data_ptr()
            else:
                self.patch_enc.forward(self.dummy, 1)
                patches = self.patch_enc.slot_ptr(self.patch_slot)
            g, bta = (self.ln_pre[0].data_ptr(), self.ln_pre[1].data_ptr()) if self.ln_pre is not None else (None, None)
            _lib.check(lib.pvr_vit_embed_f32(patches, self.cls.data_ptr(), self.pos.data_ptr(), n, S, W, g, bta,
                                             self.eps, self.x.data_ptr(), _lib.current_stream_ptr()),
                       "pvr_vit_embed_f32")
            for blk in self.blocks:
                self._ln(self.x, 1, M, blk["ln1"], self.y)
                self._gemm(self.y, blk["wqkv"], blk["bqkv"], self.qkv, M, 3 * W, W)
                _lib.check(lib.pvr_attention_f32(self.qkv.data_ptr(), n, S, W, self.heads, self.att.data_ptr(),
                                                 _lib.current_stream_ptr()), "pvr_attention_f32")
                self._gemm(self.att, blk["wo"], blk["bo"], self.x, M, W, W, res=self.x)
                self._ln(self.x, 1, M, blk["ln2"], self.y)
                self._gemm(self.y, blk["w1"], blk["b1"], self.h, M, 4 * W, W, act=self.act)
                self._gemm(self.h, blk["w2"], blk["b2"], self.x, M, W, 4 * W, res=self.x)
            if self.proj_t is None:  # MAE: the normalised class token is the embedding
                if ld != W:
                    raise _lib.PvrError("ViTRunnerF32: the embedding rows must be dense (ld == width)")
                self._ln(self.x, S, n, self.ln_post, out)
                return
            self._ln(self.x, S, n, self.ln_post, self.clsy)
            _lib.check(lib.pvr_gemm_f32(self.clsy.data_ptr(), W, self.proj_t.data_ptr(), self.zero_bias.data_ptr(), None,
                                        0, out.data_ptr(), ld, n, self.O, W, 0, _lib.current_stream_ptr()),
                       "pvr_gemm_f32")
