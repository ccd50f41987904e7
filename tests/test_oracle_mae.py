"""MAE encoders: the oracle restatement against the reference-generated goldens (tests/golden/mae.npz, written by
oracle/make_golden.py from the unmodified reference on top of the timm-0.5.4 restatement), and the host-side drop-in
container against the reference's keys / initialisation. CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, restate_mae
from pvr_habitat_b200.embeddings import EmbeddingNet, Transforms
from pvr_habitat_b200.vision_models import mae
from pvr_habitat_b200.vision_models.moco import allow_random_init


@pytest.fixture(scope="module")
def gold(golden_dir):
    """mae.npz (mae_base / mae_large) + mae_huge.npz (the first frame of each frame set), one key space."""
    g = dict(np.load(os.path.join(golden_dir, "mae.npz")))
    huge = np.load(os.path.join(golden_dir, "mae_huge.npz"))
    assert np.array_equal(huge["frames224"], g["frames224"][:1]) and np.array_equal(huge["frames64"], g["frames64"][:1])
    g.update({k: huge[k] for k in huge.files if k.endswith("mae_huge")})
    g["files"] = [k for k in g]
    return g


@pytest.mark.parametrize("case", ["structured_64", "structured_224", "structured_96x128", "noise_224", "adversarial_64"])
def test_bicubic_resize_crop_bit_exact_vs_reference(gold, case):
    x = np.ascontiguousarray(np.transpose(gold["in_" + case], (0, 3, 1, 2)))
    got = restate.resize_crop_u8(x, interpolation="bicubic")
    assert np.array_equal(got, gold["u8_" + case]), f"{int((got != gold['u8_' + case]).sum())} pixels differ"


@pytest.mark.parametrize("name", ["mae_base", "mae_large", "mae_huge"])
def test_oracle_embedding_matches_reference(gold, name):
    """The reference shuffles the patch tokens (random_masking, ratio 0); the restatement does not: same class token
    up to float32 summation order."""
    sd = restate_mae.mae_state(name, int(gold[f"seed_{name}"]))
    assert int(gold[f"out_size_{name}"]) == restate_mae.CONFIGS[name]["dim"]
    for tag in ("64", "224"):
        ref = gold[f"emb{tag}_{name}"]
        got = restate_mae.embedding_forward(sd, name, gold["frames" + tag][:len(ref)])
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-5
        np.testing.assert_allclose(got, ref, atol=2e-4)


def test_parameter_container_has_reference_keys_and_initialisation(gold):
    torch.manual_seed(9)
    m = mae.MAEParams("mae_base")
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(gold["init_keys"])
    for k in [f[5:] for f in gold["files"] if f.startswith("init_") and f != "init_keys"]:
        t = sd[k].reshape(-1)
        got = t[:: max(1, t.numel() // 64)][:64].numpy()
        np.testing.assert_allclose(got, gold["init_" + k], atol=1e-6, err_msg=k)
    assert torch.equal(mae.sincos_pos_embed(768, 14), restate_mae.sincos_table(768, 14))
    # encoder checkpoints (mae_pretrain_vit_*.pth hold the encoder only) load with strict=False like the reference
    missing = m.load_state_dict(restate_mae.mae_state("mae_base", 1), strict=False)
    assert not missing.unexpected_keys and all(k.startswith(("decoder_", "mask_token")) for k in missing.missing_keys)


def test_embedding_net_surface():
    with allow_random_init():
        net = EmbeddingNet("mae_large", disable_cuda=True)
    assert net.out_size == 1024 and tuple(net.in_shape) == (3, 224, 224) and not net.training
    assert isinstance(net.transforms, Transforms) and net.transforms.interpolation == "bicubic"
    assert all(k.startswith("embedding.") for k in net.state_dict())
    with pytest.raises(FileNotFoundError):
        EmbeddingNet("mae_base", disable_cuda=True)  # the reference's torch.load of the hard-coded path fails the same way
    with allow_random_init():
        huge = EmbeddingNet("mae_huge", disable_cuda=True)  # patch 14: 257 tokens, 16 heads of 80 (mae.py:291-296)
    assert huge.out_size == 1280 and huge.embedding.patch_size == 14 and len(huge.embedding.blocks) == 32
    with pytest.raises(Exception):
        net(torch.zeros(1, 224, 224, 3, dtype=torch.uint8))  # no CPU fallback
