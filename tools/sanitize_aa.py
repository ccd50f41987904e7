"""Small driver for compute-sanitizer: the antialiased bicubic preprocessing kernel on an up-scaling and a down-scaling
geometry, all three output formats. Usage: compute-sanitizer python tools/sanitize_aa.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvr_habitat_b200 import _lib  # noqa: E402
from pvr_habitat_b200.embeddings import CLIP_MEAN, CLIP_STD, Transforms  # noqa: E402

t = Transforms(CLIP_MEAN, CLIP_STD, size=224, crop=224, interpolation="bicubic_aa")
for (h, w, nf, n) in ((64, 64, 2, 3), (480, 640, 1, 2), (100, 75, 1, 1)):
    obs = torch.from_numpy(np.random.default_rng(h).integers(0, 256, (n, h, w, 3 * nf), dtype=np.uint8)).cuda()
    f32 = torch.empty(nf * n, 3, 224, 224, device="cuda")
    t.run(obs, nf, f32.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
    bf = torch.empty(nf * n, 224, 224, 4, dtype=torch.bfloat16, device="cuda")
    t.run(obs, nf, bf.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    f4 = torch.empty(nf * n, 224, 224, 4, device="cuda")
    t.run(obs, nf, f4.data_ptr(), _lib.PVR_FMT_NHWC4_F32, True)
    torch.cuda.synchronize()
    print((h, w, nf, n), "ok", float(f32.abs().max()), bool(torch.isfinite(f32).all()))
