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
