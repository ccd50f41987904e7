"""-m gpu: the antialiased bicubic preprocessing kernel (pvr_preprocess_u8_aa, CLIP transforms for frames that are not
224x224, src/embeddings.py:309-314) against the oracle, which is bit-identical with ATen, and against the reference's
own transforms (tests/golden/clip_transforms.npz). First run on B200 in round 2: 14 / 14 bit-exact, compute-sanitizer
clean (profiles/r02_aa_first_gpu_run.txt)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, restate_vit
from pvr_habitat_b200 import _lib
from pvr_habitat_b200.embeddings import CLIP_MEAN, CLIP_STD, Transforms

pytestmark = [pytest.mark.gpu]


def cuda_clip_transforms(obs_nhwc, nf=1):
    t = Transforms(CLIP_MEAN, CLIP_STD, size=224, crop=224, interpolation="bicubic_aa")
    n = obs_nhwc.shape[0]
    out = torch.full((nf * n, 3, 224, 224), float("nan"), device="cuda")
    t.run(torch.from_numpy(obs_nhwc).cuda(), nf, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("hw,nf,n", [((64, 64), 1, 3), ((64, 64), 2, 2), ((96, 128), 1, 2), ((100, 75), 1, 2),
                                     ((480, 640), 1, 1), ((300, 200), 3, 1), ((224, 224), 1, 2)])
def test_clip_preprocess_bit_exact_vs_oracle(hw, nf, n):
    obs = np.random.default_rng(hw[0] * 5 + hw[1] + nf).integers(0, 256, (n, hw[0], hw[1], 3 * nf), dtype=np.uint8)
    frames, _ = restate.split_frames(obs)
    want = restate_vit.clip_transforms(frames)
    got = cuda_clip_transforms(obs, nf)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{int((got != want).sum())} values differ"


def test_clip_preprocess_bf16_layout_sample_major():
    obs = restate.structured_frames(3, 64, 64, 6, 9)
    t = Transforms(CLIP_MEAN, CLIP_STD, size=224, crop=224, interpolation="bicubic_aa")
    out = torch.zeros(6, 224, 224, 4, dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, out.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    f32 = cuda_clip_transforms(obs, 2)  # frame-major
    want = torch.from_numpy(f32).reshape(2, 3, 3, 224, 224).permute(1, 0, 3, 4, 2).reshape(6, 224, 224, 3)
    assert torch.equal(out[..., :3].cpu(), want.to(torch.bfloat16)) and float(out[..., 3].abs().max()) == 0


@pytest.mark.parametrize("case", ["structured_64", "structured_96x128", "noise_100x75", "structured_336x448",
                                  "adversarial_64", "noise_224"])
def test_clip_preprocess_bit_exact_vs_reference_golden(golden_dir, case):
    """Against the uint8 images / normalisation table of the reference's own 'clip_vit' transforms
    (tests/golden/clip_transforms.npz)."""
    g = np.load(os.path.join(golden_dir, "clip_transforms.npz"))
    u = g["u8_" + case]
    want = np.stack([g["lut"][c][u[:, c]] for c in range(3)], 1)
    got = cuda_clip_transforms(g["in_" + case])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{int((got != want).sum())} values differ"
