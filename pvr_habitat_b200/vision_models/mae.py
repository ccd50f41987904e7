"""MAE ViT image encoders (`mae_base`, `mae_large`) — drop-in for src/vision_models/mae.py as the reference uses it:
`_get_embedding` builds `mae_vit_{base,large}_patch16()`, loads `checkpoint['model']` with strict=False
(src/embeddings.py:137-144) and `EmbeddingNet._forward` calls `forward_encoder(x, mask_ratio=0.0)[0][:, 0, :]`
(src/embeddings.py:377-379): the class token after the final LayerNorm, width 768 / 1024.

The parameter container keeps the reference's module tree and names (`patch_embed.proj`, `cls_token`, `pos_embed`,
`blocks.N.{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}`, `norm`, and the decoder the encoder never runs), built
from stock nn.Linear / nn.LayerNorm / nn.Conv2d in the reference's construction order and re-initialised in its order
(mae.py:117-146), so the same torch seed gives the same weights and reference state_dicts load strictly. The blocks are
timm 0.5.4 `Block`s (timm/models/vision_transformer.py: pre-LN, fused qkv Linear with bias, erf GELU, LayerNorm eps 1e-6
from mae.py:278); timm is not installed, the arithmetic is restated in oracle/restate_mae.py. The forward runs in
libpvr_b200 through clip_vit.ViTRunner (tcgen05 GEMMs and attention); there is no torch forward.

Difference kept out of the arithmetic: `random_masking` with ratio 0 (mae.py:175-200) still shuffles the patch tokens
with `torch.rand`; the class-token output is invariant to the token order (the positional embedding is added before
the shuffle), so the encoder here runs the tokens in raster order and draws no random numbers.
"""
import os

import numpy as np
import torch
from torch import nn

from .. import _lib
from .clip_vit import ViTRunner, ViTRunnerF32
from .moco import _ALLOW_RANDOM_INIT

_CONFIGS = {  # mae.py:275-298
    "mae_base": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, checkpoint="mae_pretrain_vit_base.pth"),
    "mae_large": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16, checkpoint="mae_pretrain_vit_large.pth"),
    # 16 x 16 patches of 14 x 14 + class token = 257 tokens, 16 heads of 80: attention_mma.cu, vit_patchify.cu
    "mae_huge": dict(patch_size=14, embed_dim=1280, depth=32, num_heads=16, checkpoint="mae_pretrain_vit_huge.pth"),
}


def sincos_pos_embed(dim, grid_size):
    """The fixed 2-D sin-cos table of mae.py:23-69 with the class-token row of zeros in front: for patch (row, col) the
    first dim/2 values encode `col`, the last dim/2 encode `row`, each as [sin(pos * w) | cos(pos * w)] with
    w_i = 10000^(-i / (dim/4)). float64 arithmetic like the reference's numpy code, returned (1, 1 + grid², dim) float32."""
    quarter = dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    rows, cols = np.divmod(np.arange(grid_size * grid_size), grid_size)

    def axis(pos):
        ang = pos.astype(np.float32).astype(np.float64)[:, None] * omega[None, :]
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)

    table = np.concatenate([axis(cols), axis(rows)], axis=1)
    table = np.concatenate([np.zeros((1, dim)), table], axis=0)
    return torch.from_numpy(table).float().unsqueeze(0)


class _PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, 4 * dim)


class MAEParams(nn.Module):
    """Parameter container with MaskedAutoencoderViT's state_dict keys (mae.py:75-113)."""

    def __init__(self, name, img_size=224, decoder_embed_dim=512, decoder_depth=8):
        super().__init__()
        cfg = _CONFIGS[name]
        self.name, self.img_size, self.patch_size = name, img_size, cfg["patch_size"]
        self.embed_dim, self.depth, self.num_heads = cfg["embed_dim"], cfg["depth"], cfg["num_heads"]
        dim, grid = self.embed_dim, img_size // self.patch_size
        self.patch_embed = _PatchEmbed(self.patch_size, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, dim), requires_grad=False)
        self.blocks = nn.ModuleList([_Block(dim) for _ in range(self.depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        # decoder: never executed by forward_encoder; kept so that state_dicts interchange with the reference
        self.decoder_embed = nn.Linear(dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([_Block(decoder_embed_dim) for _ in range(decoder_depth)])
        self.decoder_norm = nn.LayerNorm(decoder_embed_dim, eps=1e-6)
        self.decoder_pred = nn.Linear(decoder_embed_dim, self.patch_size ** 2 * 3, bias=True)
        self.out_size = dim
        self._runner = None
        self._initialize(grid)

    def _initialize(self, grid):
        """mae.py:117-146, same order of random draws."""
        self.pos_embed.data.copy_(sincos_pos_embed(self.embed_dim, grid))
        self.decoder_pos_embed.data.copy_(sincos_pos_embed(self.decoder_pos_embed.shape[-1], grid))
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.normal_(self.cls_token, std=.02)
        nn.init.normal_(self.mask_token, std=.02)
        for m in self.modules():  # nn.Module.apply visits children before the module itself; only leaves matter here
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def forward(self, *a, **k):
        raise _lib.PvrError("MAEParams holds parameters only; use EmbeddingNet (CUDA program), no torch fallback")

    def invalidate(self):
        self._runner = None

    def runner(self, device, precision='bf16'):
        cls = ViTRunnerF32 if precision == 'fp32' else ViTRunner
        if type(self._runner) is not cls or self._runner.device != torch.device(device):
            self._runner = cls(mae_spec(self), device)
        return self._runner


def mae_spec(m):
    blocks = [dict(ln1=(b.norm1.weight, b.norm1.bias), ln2=(b.norm2.weight, b.norm2.bias),
                   wqkv=b.attn.qkv.weight, bqkv=b.attn.qkv.bias, wo=b.attn.proj.weight, bo=b.attn.proj.bias,
                   w1=b.mlp.fc1.weight, b1=b.mlp.fc1.bias, w2=b.mlp.fc2.weight, b2=b.mlp.fc2.bias) for b in m.blocks]
    return dict(patch=m.patch_size, width=m.embed_dim, heads=m.num_heads, resolution=m.img_size,
                patch_weight=m.patch_embed.proj.weight, patch_bias=m.patch_embed.proj.bias, cls=m.cls_token,
                pos=m.pos_embed, ln_pre=None, blocks=blocks, ln_post=(m.norm.weight, m.norm.bias), proj=None,
                eps=1e-6, act=3)


def load(name, checkpoint_path=None):
    """`mae_vit_*_patch1x()` + `load_state_dict(torch.load(path)['model'], strict=False)` of src/embeddings.py:137-148."""
    if name not in _CONFIGS:
        raise NotImplementedError("Requested model not available.")
    model = MAEParams(name)
    path = checkpoint_path or _CONFIGS[name]["checkpoint"]
    if os.path.isfile(path):
        model.load_state_dict(torch.load(path, map_location="cpu")["model"], strict=False)
    elif not _ALLOW_RANDOM_INIT[-1]:
        raise FileNotFoundError(f"No such file or directory: '{path}'")
    return model
