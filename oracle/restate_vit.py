"""ORACLE — test infrastructure only. CPU restatement (torch fp32) of the CLIP ViT-B image encoder the reference reaches
through `clip.load("ViT-B/32")[0].encode_image` (src/embeddings.py:303-304, 375-376).

The arithmetic lives in openai/CLIP (`clip/model.py`, unpinned HEAD in requirements.txt:19), which is neither vendored
in the reference nor installed here. Published algorithm restated below (VisionTransformer.forward / ResidualAttention
Block / QuickGELU); the restatement is pinned against the independent implementation in `transformers`
(CLIPVisionModelWithProjection, hf:models/clip/modeling_clip.py) by tests/test_oracle_vit.py — parity with openai/CLIP
itself is unpinned (no copy of it is available offline).
State-dict keys are openai/CLIP's (`visual.*`).
"""
import numpy as np
import torch
import torch.nn.functional as F

CLIP_MEAN, CLIP_STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
W, HEADS, LAYERS, OUT = 768, 12, 12, 512


def vit_forward(sd, x):
    """x (N,3,224,224) float32 normalised frames -> (N, 512) image embeddings."""
    p = "visual."
    x = F.conv2d(x, sd[p + "conv1.weight"], stride=sd[p + "conv1.weight"].shape[-1])  # (N, W, g, g), no bias
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                          # (N, g*g, W)
    cls = sd[p + "class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + sd[p + "positional_embedding"]
    x = F.layer_norm(x, (W,), sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], 1e-5)
    n, s, _ = x.shape
    for i in range(LAYERS):
        b = f"{p}transformer.resblocks.{i}."
        y = F.layer_norm(x, (W,), sd[b + "ln_1.weight"], sd[b + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"])  # nn.MultiheadAttention
        q, k, v = (t.reshape(n, s, HEADS, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
        att = torch.softmax((q * 64 ** -0.5) @ k.transpose(-1, -2), -1) @ v
        att = att.transpose(1, 2).reshape(n, s, W)
        x = x + F.linear(att, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])
        y = F.layer_norm(x, (W,), sd[b + "ln_2.weight"], sd[b + "ln_2.bias"], 1e-5)
        y = F.linear(y, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])
        y = y * torch.sigmoid(1.702 * y)  # QuickGELU
        x = x + F.linear(y, sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
    x = F.layer_norm(x[:, 0, :], (W,), sd[p + "ln_post.weight"], sd[p + "ln_post.bias"], 1e-5)
    return x @ sd[p + "proj"]


def clip_transforms(frames_nhwc_u8):
    """src/embeddings.py:309-314: Resize(224, BICUBIC, antialias=True) -> CenterCrop(224) (the identity for 224x224
    frames, ATen's separable antialiased bicubic otherwise: oracle/restate.py:resize_bicubic_aa_f32), then /255 and
    CLIP's Normalize (three rounded fp32 operations, as in oracle/restate.py:normalize_lut)."""
    from oracle import restate
    lut = restate.normalize_lut(CLIP_MEAN, CLIP_STD)
    u = np.ascontiguousarray(np.transpose(frames_nhwc_u8, (0, 3, 1, 2)))
    u = restate.resize_crop_u8(u, 224, 224, interpolation="bicubic_aa")
    return np.stack([lut[c][u[:, c]] for c in range(3)], 1)


def embedding_forward(sd, frames_nhwc_u8):
    with torch.no_grad():
        return vit_forward(sd, torch.from_numpy(clip_transforms(frames_nhwc_u8))).numpy()


def vit_state(patch, seed, gain=1.0):
    """Deterministic (numpy default_rng) weights with openai/CLIP key names, drawn with the standard deviations of
    CLIP.initialize_parameters (x `gain` for the transformer matrices; gain 2 is used as a stress case); non-trivial
    LayerNorm affine and biases so every term of the forward is exercised."""
    rng = np.random.default_rng(seed)
    t = lambda *s, std=1.0: torch.from_numpy((rng.standard_normal(s) * std).astype(np.float32))  # noqa: E731
    g = 224 // patch
    sd = {"visual.conv1.weight": t(W, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
          "visual.class_embedding": t(W, std=W ** -0.5),
          "visual.positional_embedding": t(g * g + 1, W, std=W ** -0.5 * 4),
          "visual.proj": t(W, OUT, std=W ** -0.5)}
    for name in ("ln_pre", "ln_post"):
        sd[f"visual.{name}.weight"] = 1 + t(W, std=0.1)
        sd[f"visual.{name}.bias"] = t(W, std=0.1)
    proj_std, attn_std, fc_std = (W ** -0.5) * ((2 * LAYERS) ** -0.5), W ** -0.5, (2 * W) ** -0.5
    for i in range(LAYERS):
        b = f"visual.transformer.resblocks.{i}."
        sd[b + "attn.in_proj_weight"] = t(3 * W, W, std=attn_std * gain)
        sd[b + "attn.in_proj_bias"] = t(3 * W, std=0.02)
        sd[b + "attn.out_proj.weight"] = t(W, W, std=proj_std * gain)
        sd[b + "attn.out_proj.bias"] = t(W, std=0.02)
        sd[b + "mlp.c_fc.weight"] = t(4 * W, W, std=fc_std * gain)
        sd[b + "mlp.c_fc.bias"] = t(4 * W, std=0.02)
        sd[b + "mlp.c_proj.weight"] = t(W, 4 * W, std=proj_std * gain)
        sd[b + "mlp.c_proj.bias"] = t(W, std=0.02)
        for ln in ("ln_1", "ln_2"):
            sd[b + ln + ".weight"] = 1 + t(W, std=0.1)
            sd[b + ln + ".bias"] = t(W, std=0.1)
    return sd


def to_hf_state(sd):
    """Key mapping openai/CLIP -> transformers CLIPVisionModelWithProjection (for the pin test)."""
    out = {"vision_model.embeddings.class_embedding": sd["visual.class_embedding"],
           "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
           "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
           "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"],
           "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"],
           "vision_model.post_layernorm.weight": sd["visual.ln_post.weight"],
           "vision_model.post_layernorm.bias": sd["visual.ln_post.bias"],
           "visual_projection.weight": sd["visual.proj"].t().contiguous()}
    for i in range(LAYERS):
        b, h = f"visual.transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[b + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = sd[b + "attn.in_proj_bias"].chunk(3, 0)
        for n, w_, b_ in (("q", wq, bq), ("k", wk, bk), ("v", wv, bv)):
            out[h + f"self_attn.{n}_proj.weight"], out[h + f"self_attn.{n}_proj.bias"] = w_, b_
        out[h + "self_attn.out_proj.weight"] = sd[b + "attn.out_proj.weight"]
        out[h + "self_attn.out_proj.bias"] = sd[b + "attn.out_proj.bias"]
        out[h + "layer_norm1.weight"], out[h + "layer_norm1.bias"] = sd[b + "ln_1.weight"], sd[b + "ln_1.bias"]
        out[h + "layer_norm2.weight"], out[h + "layer_norm2.bias"] = sd[b + "ln_2.weight"], sd[b + "ln_2.bias"]
        out[h + "mlp.fc1.weight"], out[h + "mlp.fc1.bias"] = sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"]
        out[h + "mlp.fc2.weight"], out[h + "mlp.fc2.bias"] = sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"]
    return out
