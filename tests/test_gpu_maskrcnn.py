"""-m gpu: EmbeddingNet('maskrcnn_l3') (src/embeddings.py:283-295, :380-383; src/vision_models/maskrcnn.py) through
the C ABI: the float-resize / row-permutation mode of the preprocessing kernel bit for bit against the oracle, the
stride-on-1x1 ResNet-50 program + compression head against the reference goldens (tests/golden/maskrcnn.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate_maskrcnn as rm
from pvr_habitat_b200 import _lib
from pvr_habitat_b200.embeddings import EmbeddingNet
from pvr_habitat_b200.vision_models.moco import allow_random_init

pytestmark = pytest.mark.gpu
CASES = ["structured_64", "structured_224", "small_40x48", "adversarial_64"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "maskrcnn.npz"))


@pytest.fixture(scope="module")
def net(gold):
    with allow_random_init():
        n = EmbeddingNet("maskrcnn_l3")
    n.embedding.load_state_dict(rm.maskrcnn_state(int(gold["seed"])), strict=True)
    n.invalidate()
    return n


@pytest.mark.parametrize("case", CASES)
def test_float_resize_transforms_bit_exact(gold, net, case):
    """NCHW float32 output of the kernel == the reference transforms (golden slices) == the oracle (every pixel)."""
    frames = gold["in_" + case]
    got = net.preprocess(torch.from_numpy(frames)).cpu().numpy()
    assert np.array_equal(got[:, :, :12], gold["t_top_" + case])
    assert np.array_equal(got[:, :, ::7, ::5], gold["t_sub_" + case])
    assert np.array_equal(got, rm.maskrcnn_transforms(frames))


def test_float_resize_other_output_formats(gold, net):
    """NHWC4 float32 (fp32 mode input) is the same numbers; NHWC4 bf16 their round-to-nearest; 6-channel observations
    split into two frames like the uint8 modes."""
    frames = gold["in_small_40x48"]
    want = torch.from_numpy(rm.maskrcnn_transforms(frames)).permute(0, 2, 3, 1)
    obs = torch.from_numpy(np.concatenate([frames, frames[:, ::-1].copy()], -1)).cuda()
    want2 = torch.from_numpy(rm.maskrcnn_transforms(frames[:, ::-1].copy())).permute(0, 2, 3, 1)
    out = torch.full((2, 224, 224, 4), float("nan"), device="cuda")
    net.transforms.run(obs, 2, out.data_ptr(), _lib.PVR_FMT_NHWC4_F32, False)
    torch.cuda.synchronize()
    assert torch.equal(out[0, ..., :3].cpu(), want[0]) and torch.equal(out[1, ..., :3].cpu(), want2[0])
    assert not out[..., 3].any()
    outb = torch.full((2, 224, 224, 4), float("nan"), device="cuda", dtype=torch.bfloat16)
    net.transforms.run(obs, 2, outb.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, False)
    torch.cuda.synchronize()
    assert torch.equal(outb[0, ..., :3].cpu(), want[0].bfloat16())


@pytest.mark.parametrize("case", CASES)
def test_embedding_vs_reference_golden(gold, net, case):
    """north star: bf16 embeddings within relative L2 <= 1e-2 and cosine >= 0.999 of the reference."""
    got = np.atleast_2d(net(torch.from_numpy(gold["in_" + case]))).astype(np.float64)
    ref = gold["emb_" + case].astype(np.float64)
    assert got.shape == ref.shape
    r = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    print(f"maskrcnn_l3 {case}: rel-L2 {r:.2e}, min cos {cos.min():.5f}")
    assert r <= 1e-2 and cos.min() >= 0.999


def test_embedding_fp32_mode_vs_reference_golden(gold):
    with allow_random_init():
        n = EmbeddingNet("maskrcnn_l3")
    n.embedding.load_state_dict(rm.maskrcnn_state(int(gold["seed"])), strict=True)
    n.set_precision('fp32')
    for case in ("structured_64", "small_40x48"):
        got = np.atleast_2d(n(torch.from_numpy(gold["in_" + case]))).astype(np.float64)
        ref = gold["emb_" + case].astype(np.float64)
        r = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"maskrcnn_l3 {case} fp32 mode: rel-L2 {r:.2e}")
        assert r <= 1e-5, r
