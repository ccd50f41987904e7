"""ORACLE — test infrastructure only. CPU restatement (torch fp32) of the MAE ViT encoders the reference reaches through
`EmbeddingNet('mae_base' | 'mae_large')`: src/embeddings.py:81 (bicubic Resize), :137-144 (construction / checkpoint),
:377-379 (`forward_encoder(x, mask_ratio=0.0)[0][:, 0, :]`), src/vision_models/mae.py:202-222 (forward_encoder).

Two layers of code are involved:
  * the reference's own `mae.py` (patch embedding + fixed sin-cos positional table + class token + `random_masking` +
    final norm): executed UNMODIFIED by oracle/make_golden.py -> tests/golden/mae.npz, which pins `mae_forward` below;
  * timm 0.5.4 `PatchEmbed` / `Block` (requirements.txt:20 pins "timm=0.5.4"), imported by mae.py:20. timm is neither
    vendored in the reference nor installed here: its published algorithm (timm/models/vision_transformer.py, 0.5.4:
    Attention = fused qkv Linear -> softmax(q k^T / sqrt(d)) v -> proj; Mlp = fc1 -> nn.GELU (erf) -> fc2; Block =
    x + attn(norm1(x)), x + mlp(norm2(x)); timm/models/layers/patch_embed.py: Conv2d(kernel = stride = patch) ->
    flatten(2).transpose(1, 2)) is restated in the `PatchEmbed` / `Block` classes below, which the golden generator
    hands to the reference as the `timm.models.vision_transformer` module. Parity with timm's own code is therefore
    UNPINNED (no copy of it is available offline); parity with the reference's mae.py on top of that restatement is pinned.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

CONFIGS = {"mae_base": dict(dim=768, depth=12, heads=12, patch=16), "mae_large": dict(dim=1024, depth=24, heads=16, patch=16),
           "mae_huge": dict(dim=1280, depth=32, heads=16, patch=14)}  # mae.py:275-296
CHECKPOINTS = {"mae_base": "mae_pretrain_vit_base.pth", "mae_large": "mae_pretrain_vit_large.pth",
               "mae_huge": "mae_pretrain_vit_huge.pth"}  # embeddings.py:139,143,147


# ---------------------------------------------------------------------------- timm 0.5.4, restated (see header)
class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads, self.scale = num_heads, (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ qkv[2]).transpose(1, 2).reshape(b, n, c))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


# ---------------------------------------------------------------------------- the encoder, functional
def sincos_table(dim, grid):
    """mae.py:23-69 (get_2d_sincos_pos_embed with cls_token=True): the meshgrid puts the column index first, so the
    first half of the channels encodes the patch column and the second half the row; float64, then .float()."""
    omega = 1.0 / 10000 ** (np.arange(dim // 4, dtype=np.float64) / (dim / 4.0))
    col, row = np.meshgrid(np.arange(grid, dtype=np.float32), np.arange(grid, dtype=np.float32))
    halves = []
    for pos in (col.reshape(-1), row.reshape(-1)):
        ang = np.einsum("m,d->md", pos, omega)
        halves.append(np.concatenate([np.sin(ang), np.cos(ang)], axis=1))
    table = np.concatenate([np.zeros([1, dim]), np.concatenate(halves, axis=1)], axis=0)
    return torch.from_numpy(table).float().unsqueeze(0)


def mae_forward(sd, name, x):
    """x (N,3,224,224) float32 normalised frames -> (N, dim): class token after the final norm (mae.py:202-222 with
    mask_ratio 0; the token shuffle of random_masking is skipped: the class-token output does not depend on the order
    of the patch tokens — tests/test_oracle_mae.py checks it against the reference, which does shuffle)."""
    c = CONFIGS[name]
    dim, heads = c["dim"], c["heads"]
    x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=c["patch"])
    x = x.flatten(2).transpose(1, 2) + sd["pos_embed"][:, 1:, :]
    cls = (sd["cls_token"] + sd["pos_embed"][:, :1, :]).expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], 1)
    n, s, _ = x.shape
    for i in range(c["depth"]):
        b = f"blocks.{i}."
        y = F.layer_norm(x, (dim,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"]).reshape(n, s, 3, heads, dim // heads)
        q, k, v = qkv.permute(2, 0, 3, 1, 4)
        att = torch.softmax((q @ k.transpose(-2, -1)) * (dim // heads) ** -0.5, -1) @ v
        x = x + F.linear(att.transpose(1, 2).reshape(n, s, dim), sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        y = F.layer_norm(x, (dim,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
        y = F.gelu(F.linear(y, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
        x = x + F.linear(y, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
    return F.layer_norm(x[:, 0, :], (dim,), sd["norm.weight"], sd["norm.bias"], 1e-6)


def mae_transforms(frames_nhwc_u8):
    """src/embeddings.py:80-85 with interpolation=3: bicubic Resize(256) -> CenterCrop(224) -> /255 -> Normalize."""
    from oracle import restate
    return restate.transforms(np.ascontiguousarray(np.transpose(frames_nhwc_u8, (0, 3, 1, 2))), interpolation="bicubic")


def embedding_forward(sd, name, frames_nhwc_u8):
    with torch.no_grad():
        return mae_forward(sd, name, torch.from_numpy(mae_transforms(frames_nhwc_u8))).numpy()


def mae_state(name, seed):
    """Deterministic (numpy default_rng: platform independent) encoder weights under the checkpoint's key names
    (`checkpoint['model']` of mae_pretrain_vit_*.pth holds the encoder only). Xavier-uniform matrices like
    mae.py:137-146, but non-trivial LayerNorm affines and biases so that every term of the forward is exercised."""
    c = CONFIGS[name]
    dim, grid = c["dim"], 224 // c["patch"]
    rng = np.random.default_rng(seed)

    def xavier(o, i, *rest):
        fan_in, fan_out = i * int(np.prod(rest or (1,))), o * int(np.prod(rest or (1,)))
        a = (6.0 / (fan_in + fan_out)) ** 0.5
        return torch.from_numpy(rng.uniform(-a, a, (o, i) + tuple(rest)).astype(np.float32))

    def normal(*s, std):
        return torch.from_numpy((rng.standard_normal(s) * std).astype(np.float32))

    sd = {"cls_token": normal(1, 1, dim, std=0.02), "pos_embed": sincos_table(dim, grid),
          "patch_embed.proj.weight": xavier(dim, 3, c["patch"], c["patch"]),
          "patch_embed.proj.bias": normal(dim, std=0.02)}
    for i in range(c["depth"]):
        b = f"blocks.{i}."
        for ln in ("norm1", "norm2"):
            sd[b + ln + ".weight"] = 1 + normal(dim, std=0.1)
            sd[b + ln + ".bias"] = normal(dim, std=0.05)
        sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"] = xavier(3 * dim, dim), normal(3 * dim, std=0.02)
        sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"] = xavier(dim, dim), normal(dim, std=0.02)
        sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"] = xavier(4 * dim, dim), normal(4 * dim, std=0.02)
        sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"] = xavier(dim, 4 * dim), normal(dim, std=0.02)
    sd["norm.weight"], sd["norm.bias"] = 1 + normal(dim, std=0.1), normal(dim, std=0.05)
    return sd
