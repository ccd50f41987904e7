"""Small invocations of the kernels added at the end of round 2, for compute-sanitizer (development tool):
    compute-sanitizer --tool memcheck  python tools/sanitize_new_kernels.py
    compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py
small_conv1_fwd / _wgrad (mma.sync, cp.async pipeline), vit_attention_mma (head_dim 80, 257 tokens, and the attention
pool's 50 tokens), vit_patchify, avgpool2 (through the clip_rn50 trunk), attnpool_tokens, the float-resize mode of the
preprocessing kernel, and the A-stationary tile order of conv_gemm_kernel (PVR_ASTAT=2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PVR_ASTAT"] = "2"
from pvr_habitat_b200 import _lib  # noqa: E402
from pvr_habitat_b200.embeddings import EmbeddingNet  # noqa: E402
from pvr_habitat_b200.vision_models.moco import allow_random_init  # noqa: E402

lib, st = _lib.lib(), _lib.current_stream_ptr
g = torch.Generator(device="cuda").manual_seed(1)

# first small-conv layer: forward through the 'random' encoder (ragged frame count), weight gradient directly
with allow_random_init():
    net = EmbeddingNet("random")
print("random", net(torch.randint(0, 256, (5, 64, 64, 3), dtype=torch.uint8)).shape)
F, H = 3, 30
Ho = (H - 1) // 2 + 1
x = torch.zeros(F, H, H, 4, device="cuda")
x[..., :3] = torch.rand(F, H, H, 3, device="cuda", generator=g)
y = (torch.randn(F, Ho, Ho, 32, device="cuda", generator=g) * 0.7).bfloat16()
dy = torch.randn(F * Ho * Ho, 32, device="cuda", generator=g)
dw, db = torch.zeros(32, 3, 3, 4, device="cuda"), torch.zeros(32, device="cuda")
_lib.check(lib.pvr_small_conv1_wgrad(dy.data_ptr(), y.data_ptr(), 32, x.bfloat16().data_ptr(), F, H, H, Ho, Ho,
                                     dw.data_ptr(), db.data_ptr(), st()), "wgrad")
print("wgrad", float(dw.abs().sum()))

# attention: mae_huge's shape and the attention pool's
for tokens, n, heads, d in ((257, 2, 16, 80), (50, 3, 32, 64), (33, 1, 2, 128)):
    w = heads * d
    qkv = torch.randn(n * tokens, 3 * w, device="cuda", generator=g).bfloat16()
    out = torch.empty(n * tokens, w, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.pvr_attention_mma(qkv.data_ptr(), n, tokens, w, heads, out.data_ptr(), st()), "attention_mma")
    print("attention", tokens, d, float(out.float().abs().mean()))

# patch gather (p = 14), attention-pool tokens
fr = torch.randn(2, 224, 224, 4, device="cuda", generator=g).bfloat16()
col = torch.empty(2 * 256, 640, dtype=torch.bfloat16, device="cuda")
_lib.check(lib.pvr_vit_patchify(fr.data_ptr(), 2, 224, 14, 640, 0, col.data_ptr(), st()), "patchify")
xt = torch.randn(2, 49, 2048, device="cuda", generator=g).bfloat16()
pos = torch.randn(50, 2048, device="cuda", generator=g)
tok = torch.empty(2, 50, 2048, dtype=torch.bfloat16, device="cuda")
_lib.check(lib.pvr_attnpool_tokens(xt.data_ptr(), 2, 49, 2048, pos.data_ptr(), 0, tok.data_ptr(), st()), "tokens")

# whole encoders: clip_rn50 (avgpool2 op, pooled projection blocks), maskrcnn_l3 (float resize, stride on the 1x1),
# moco uber with the A-stationary order forced
for name, n, hw in (("clip_rn50", 2, 64), ("maskrcnn_l3", 2, 40), ("moco_aug_uber_34", 16, 64)):
    with allow_random_init():
        net = EmbeddingNet(name)
    e = net(torch.randint(0, 256, (n, hw, hw + 8, 3), dtype=torch.uint8))
    torch.cuda.synchronize()
    print(name, e.shape)
print("done")
