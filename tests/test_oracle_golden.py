"""Pins the oracle restatement (oracle/restate.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py). CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import restate


@pytest.fixture(scope="module")
def tf(golden_dir):
    return np.load(os.path.join(golden_dir, "transforms.npz"))


@pytest.fixture(scope="module")
def emb(golden_dir):
    return np.load(os.path.join(golden_dir, "embeddings.npz"))


TRANSFORM_CASES = ["structured_64", "structured_224", "structured_96x128", "noise_64", "noise_224", "adversarial_64",
                   "adversarial_224"]


@pytest.mark.parametrize("case", TRANSFORM_CASES)
def test_resize_crop_bit_exact(tf, case):
    frames = tf["in_" + case]
    got = restate.resize_crop_u8(np.ascontiguousarray(frames.transpose(0, 3, 1, 2)))
    ref = tf["u8_" + case]
    assert got.dtype == np.uint8 and got.shape == ref.shape
    if case == "structured_96x128":
        # non-dyadic ratio 128 -> 341: the installed torchvision (antialias=True default) and the reference's pin
        # (torchvision 0.10, no antialias) disagree on exact .5 ties; the oracle follows the pin (see next test)
        d = np.abs(got.astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-4
        return
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {ref.size} pixels differ"


@pytest.mark.parametrize("hw", [(96, 128), (100, 75), (480, 640), (84, 84), (210, 160), (64, 64), (224, 224)])
def test_resize_matches_aten_nonantialiased_bitwise(hw):
    """torchvision-0.10 semantics = F.interpolate(bilinear, align_corners=False) without antialias, float bits."""
    h, w = hw
    x = np.random.default_rng(h * 1000 + w).integers(0, 256, (2, 3, h, w), dtype=np.uint8)
    rh, rw, _, _ = restate.resize_geometry(h, w)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).float(), size=(rh, rw), mode="bilinear",
                                          align_corners=False).numpy()
    got = restate.resize_bilinear_f32(x, rh, rw)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("hw", [(96, 128), (100, 75), (480, 640), (84, 84), (33, 47), (64, 64), (224, 224)])
def test_bicubic_resize_matches_aten_bitwise(hw):
    """MAE transforms (src/embeddings.py:81, interpolation=3): torchvision-0.10 semantics = F.interpolate(bicubic,
    align_corners=False) without antialias (A = -0.75), float bits; then clamp + half-even round like torchvision."""
    h, w = hw
    x = np.random.default_rng(h * 1000 + w + 7).integers(0, 256, (2, 3, h, w), dtype=np.uint8)
    rh, rw, top, left = restate.resize_geometry(h, w)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).float(), size=(rh, rw), mode="bicubic",
                                          align_corners=False)
    got = restate.resize_bicubic_f32(x, rh, rw)
    assert np.array_equal(got.view(np.uint32), ref.numpy().view(np.uint32))
    u8 = torch.round(ref.clamp(min=0, max=255)).to(torch.uint8)[:, :, top:top + 224, left:left + 224].numpy()
    assert np.array_equal(restate.resize_crop_u8(x, interpolation="bicubic"), u8)
    assert (ref.numpy() < 0).any() and (ref.numpy() > 255).any()  # the clamp is exercised


@pytest.mark.parametrize("hw", [(64, 64), (96, 128), (100, 75), (480, 640), (300, 200), (84, 84), (225, 231)])
def test_antialiased_bicubic_resize_matches_aten_bitwise(hw):
    """CLIP transforms (src/embeddings.py:309-310): Resize(224, BICUBIC, antialias=True) = ATen's separable antialiased
    bicubic (a = -0.5) on the float image, up- and down-scaling, float bits; then clamp + half-even round."""
    h, w = hw
    x = np.random.default_rng(h * 1000 + w + 3).integers(0, 256, (2, 3, h, w), dtype=np.uint8)
    rh, rw, top, left = restate.resize_geometry(h, w, 224, 224)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).float(), size=(rh, rw), mode="bicubic",
                                          align_corners=False, antialias=True)
    got = restate.resize_bicubic_aa_f32(x, rh, rw)
    assert np.array_equal(got.view(np.uint32), ref.numpy().view(np.uint32))
    u8 = torch.round(ref.clamp(min=0, max=255)).to(torch.uint8)[:, :, top:top + 224, left:left + 224].numpy()
    assert np.array_equal(restate.resize_crop_u8(x, 224, 224, interpolation="bicubic_aa"), u8)


def test_antialiased_bicubic_is_the_identity_at_the_target_size():
    x = np.random.default_rng(5).integers(0, 256, (1, 3, 224, 224), dtype=np.uint8)
    assert np.array_equal(restate.resize_crop_u8(x, 224, 224, interpolation="bicubic_aa"), x)


def test_normalize_lut_bit_exact(tf):
    assert np.array_equal(restate.normalize_lut().view(np.uint32), tf["lut"].view(np.uint32))


def test_ties_are_present(tf):
    """The adversarial set must actually exercise half-to-even rounding."""
    frames = tf["in_adversarial_64"].transpose(0, 3, 1, 2).astype(np.float32)
    x = torch.nn.functional.interpolate(torch.from_numpy(frames), size=(256, 256), mode="bilinear",
                                        align_corners=False)
    frac = (x - torch.floor(x)).numpy()
    assert (frac == 0.5).sum() > 1000


def test_frame_split_regroup_roundtrip():
    rng = np.random.default_rng(0)
    obs = rng.integers(0, 256, (5, 8, 8, 9), dtype=np.uint8)
    fr, n = restate.split_frames(obs)
    assert n == 3 and fr.shape == (15, 8, 8, 3)
    assert np.array_equal(fr[5 * 2 + 3], obs[3, :, :, 6:9])  # frame-major: index f*N + i
    emb = np.arange(15 * 4, dtype=np.float32).reshape(15, 4)
    g = restate.regroup_frames(emb, 3)
    assert g.shape == (5, 12) and np.array_equal(g[3, 8:12], emb[2 * 5 + 3])


def _parts(name, seeds):
    variants = {"moco_aug": ["conv5"], "moco_aug_l4": ["l4"], "moco_aug_l3": ["l3"],
                "moco_aug_uber_34": ["l3", "l4"], "moco_aug_uber_345": ["l3", "l4", "conv5"]}[name]
    seed = {"conv5": int(seeds[0]), "l4": int(seeds[1]), "l3": int(seeds[2])}
    return [(v, restate.resnet50_state(v, seed[v])) for v in variants]


@pytest.mark.parametrize("name", ["moco_aug", "moco_aug_l4", "moco_aug_l3", "moco_aug_uber_34", "moco_aug_uber_345"])
def test_embedding_restatement_matches_reference(emb, name):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    parts = _parts(name, emb["weight_seeds"])
    got = restate.embedding_forward(parts, emb["frames64"])
    ref = emb[f"emb64_{name}"]
    assert got.shape == ref.shape == (4, int(emb[f"out_size_{name}"]))
    # same fp32 library kernels, same order of operations: tight tolerance (threads may change summation order)
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-4 * float(np.abs(ref).max()))


def test_embedding_restatement_224_and_two_frame(emb):
    parts = _parts("moco_aug", emb["weight_seeds"])
    got = restate.embedding_forward(parts, emb["frames224"])
    ref = emb["emb224_moco_aug"]
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-4 * float(np.abs(ref).max()))
    got2 = restate.embed_observations(parts, emb["obs2"], batch_size=2)  # ragged last mini-batch
    ref2 = emb["emb_obs2_moco_aug"]
    assert got2.shape == ref2.shape == (3, 4096)
    np.testing.assert_allclose(got2, ref2, rtol=2e-4, atol=2e-4 * float(np.abs(ref2).max()))


def test_uber_state_dict_is_empty_in_reference(emb):
    """Reference quirk D8 (src/embeddings.py:45-53) recorded by the fixture."""
    assert int(emb["n_state_keys_moco_aug_uber_34"]) == 0
    assert int(emb["n_state_keys_moco_aug"]) == 318


@pytest.mark.parametrize("name", ["resnet18", "resnet34"])
def test_resnet_basic_restatement_matches_reference(golden_dir, name):
    """resnet18 / resnet34: the oracle restatement against the reference's EmbeddingNet outputs
    (torchvision nets with fc = Identity, src/embeddings.py:112-117) on the same synthetic weights."""
    g = np.load(os.path.join(golden_dir, "resnet_basic.npz"))
    sd = restate.resnet_basic_state(name, int(g[f"seed_{name}"]))
    for tag in ("64", "224"):
        frames = g["frames" + tag]
        x = torch.from_numpy(restate.transforms(np.ascontiguousarray(frames.transpose(0, 3, 1, 2))))
        with torch.no_grad():
            got = restate.resnet_basic_forward(sd, name, x).numpy()
        ref = g[f"emb{tag}_{name}"]
        assert got.shape == ref.shape == (frames.shape[0], 512)
        assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max()
    assert int(g[f"out_size_{name}"]) == 512
