"""Pins oracle/restate_vit.py against the independent CLIP implementation in `transformers` (CPU only)."""
import numpy as np
import pytest
import torch

from oracle import restate, restate_vit as rv


@pytest.mark.parametrize("patch", [32, 16])
def test_vit_restatement_matches_transformers_clip(patch):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                        num_attention_heads=12, image_size=224, patch_size=patch,
                                        hidden_act="quick_gelu", projection_dim=512, layer_norm_eps=1e-5)
    hf = transformers.CLIPVisionModelWithProjection(cfg).eval()
    sd = rv.vit_state(patch, 5)
    missing = hf.load_state_dict(rv.to_hf_state(sd), strict=False)
    assert not [k for k in missing.missing_keys if "position_ids" not in k] and not missing.unexpected_keys
    frames = restate.structured_frames(2, 224, 224, 3, 31)
    x = torch.from_numpy(rv.clip_transforms(frames))
    with torch.no_grad():
        ref = hf(pixel_values=x).image_embeds.numpy()
        got = rv.vit_forward(sd, x).numpy()
    assert got.shape == ref.shape == (2, 512)
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4 * float(np.abs(ref).max()))


def test_clip_transforms_are_identity_resize_plus_normalise():
    frames = restate.structured_frames(1, 224, 224, 3, 2)
    x = rv.clip_transforms(frames)
    ref = (frames.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255.0)
           - np.array(rv.CLIP_MEAN, np.float32)[None, :, None, None]) / np.array(rv.CLIP_STD, np.float32)[None, :, None, None]
    assert np.array_equal(x, ref.astype(np.float32))


# ------------------------------------------------------------------------------------------------ CLIP transforms
@pytest.mark.parametrize("case", ["structured_64", "structured_96x128", "noise_100x75", "structured_336x448",
                                  "adversarial_64", "noise_224"])
def test_clip_transforms_bit_exact_vs_reference(golden_dir, case):
    """The `transforms` the reference builds for 'clip_vit' (src/embeddings.py:309-314, antialiased bicubic Resize) run
    by oracle/make_golden.py on frames that are not 224x224: the oracle's resize + crop gives the same uint8 image and
    the same normalisation table, hence the same float tensor bit for bit."""
    import os
    from oracle import restate
    g = np.load(os.path.join(golden_dir, "clip_transforms.npz"))
    x = np.ascontiguousarray(np.transpose(g["in_" + case], (0, 3, 1, 2)))
    got = restate.resize_crop_u8(x, 224, 224, interpolation="bicubic_aa")
    assert np.array_equal(got, g["u8_" + case]), f"{int((got != g['u8_' + case]).sum())} pixels differ"
    lut = restate.normalize_lut(rv.CLIP_MEAN, rv.CLIP_STD)
    assert np.array_equal(lut.view(np.uint32), g["lut"].view(np.uint32))
    full = rv.clip_transforms(g["in_" + case])
    want = np.stack([g["lut"][c][g["u8_" + case][:, c]] for c in range(3)], 1)
    assert np.array_equal(full.view(np.uint32), want.view(np.uint32))
