"""clip_rn50 (src/embeddings.py:305-314, 375-376): the oracle restatement of openai/CLIP's ModifiedResNet against the
goldens that oracle/make_golden.py wrote by running the UNMODIFIED reference EmbeddingNet('clip_rn50') on top of it
(`clip` itself is not installed: parity with openai/CLIP's own code is unpinned), and the host side
of the drop-in (openai key names, checkpoint handling). CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import restate_clip_rn as rc
from pvr_habitat_b200.embeddings import EmbeddingNet
from pvr_habitat_b200.vision_models import clip_rn
from pvr_habitat_b200.vision_models.moco import allow_random_init

CASES = ["structured_64", "structured_224", "structured_96x128"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "clip_rn50.npz"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_embedding_matches_reference(gold, case):
    """Same model object, but the oracle's own restatement of the transforms (antialiased bicubic, CLIP mean / std)."""
    got = rc.embedding_forward(rc.clip_rn50_state(int(gold["seed"])), gold["in_" + case])
    ref = gold["emb_" + case]
    assert got.shape == ref.shape and ref.shape[1] == 1024
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6


def test_attention_pool_is_single_query_attention():
    """F.multi_head_attention_forward as CLIP calls it == softmax((q W_q + b_q) / 8 . K^T) V for the mean token."""
    torch.manual_seed(0)
    ap = rc.AttentionPool2d(7, 2048, 32, 1024).eval()
    x = torch.randn(2, 2048, 7, 7)
    with torch.no_grad():
        want = ap(x)
        t = x.flatten(2).permute(0, 2, 1)
        t = torch.cat([t.mean(1, keepdim=True), t], 1) + ap.positional_embedding
        q = ap.q_proj(t[:, :1]).reshape(2, 1, 32, 64).transpose(1, 2) * 64 ** -0.5
        k = ap.k_proj(t).reshape(2, 50, 32, 64).transpose(1, 2)
        v = ap.v_proj(t).reshape(2, 50, 32, 64).transpose(1, 2)
        o = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(2, 2048)
        got = ap.c_proj(o)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)


def test_container_keys_and_surface(gold, tmp_path):
    with allow_random_init():
        net = EmbeddingNet("clip_rn50", disable_cuda=True)
    assert net.out_size == int(gold["out_size"]) == 1024 and tuple(net.in_shape) == (3, 224, 224) and not net.training
    assert sorted(net.state_dict().keys()) == list(gold["visual_keys"])
    assert net.transforms.interpolation == "bicubic_aa" and net.embedding.visual.input_resolution == 224
    with pytest.raises(FileNotFoundError):
        EmbeddingNet("clip_rn50", disable_cuda=True)
    sd = rc.clip_rn50_state(2)
    full = dict(sd)
    full["transformer.resblocks.0.attn.in_proj_weight"] = torch.zeros(4, 4)  # text tower: dropped by the loader
    path = str(tmp_path / "RN50.pt")
    torch.save(full, path)
    m, _ = clip_rn.load("RN50", checkpoint_path=path)
    assert all(torch.equal(m.state_dict()[k], v) for k, v in sd.items())
    # a reference checkpoint (`embedding_model_state_dict`) carries the text tower under `embedding.`: dropped, strict
    with allow_random_init():
        net = EmbeddingNet("clip_rn50", disable_cuda=True)
    ck = {"embedding." + k: v for k, v in sd.items()}
    ck["embedding.token_embedding.weight"] = torch.zeros(3, 3)
    ck["embedding.logit_scale"] = torch.zeros(())
    net.load_state_dict(ck)
    assert torch.equal(net.embedding.visual.layer4[2].conv3.weight, sd["visual.layer4.2.conv3.weight"])
