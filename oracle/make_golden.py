"""ORACLE (build container only) — generate tests/golden/*.npz by running the UNMODIFIED reference.

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

Every fixture stores the seeds / inputs and the reference's outputs. Weights are NOT stored: they are regenerated
from `oracle.restate.resnet50_state(variant, seed)` (numpy default_rng: platform independent), written here as
synthetic checkpoints under the file names the reference hard-codes, and loaded by the reference's own constructors.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, restate  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEEDS = {"moco_aug": 101, "moco_aug_l4": 102, "moco_aug_l3": 103}
VARIANT = {"moco_aug": "conv5", "moco_aug_l4": "l4", "moco_aug_l3": "l3"}


def golden_transforms(E):
    """The reference's `transforms` (src/embeddings.py:80-85), captured as the uint8 image after Resize+CenterCrop
    and the 3x256 table of ConvertImageDtype+Normalize: together they determine the float output bit for bit."""
    _, tf = E._get_embedding("random")
    resize_crop = torch.nn.Sequential(tf[0], tf[1])
    to_float = torch.nn.Sequential(tf[2], tf[3])
    out = {}
    cases = {
        "structured_64": restate.structured_frames(3, 64, 64, 3, 11),
        "structured_224": restate.structured_frames(2, 224, 224, 3, 12),
        "structured_96x128": restate.structured_frames(2, 96, 128, 3, 13),
        "noise_64": np.random.default_rng(14).integers(0, 256, (2, 64, 64, 3), dtype=np.uint8),
        "noise_224": np.random.default_rng(15).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8),
        "adversarial_64": restate.adversarial_frames(64, 64),
        "adversarial_224": restate.adversarial_frames(224, 224),
    }
    for name, frames in cases.items():
        x = torch.from_numpy(frames).permute(0, 3, 1, 2).contiguous()
        u = resize_crop(x)
        assert u.dtype == torch.uint8
        out["in_" + name] = frames
        out["u8_" + name] = u.numpy()
        # full float output cross-checked against LUT composition right here
        full = to_float(u).numpy()
        ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).repeat(1, 3, 1, 1)
        lut = to_float(ramp).numpy().reshape(3, 256)
        comp = np.stack([lut[c][u.numpy()[:, c]] for c in range(3)], 1)
        assert np.array_equal(full, comp)
        out["lut"] = lut
    np.savez_compressed(os.path.join(GOLDEN, "transforms.npz"), **out)
    print("transforms.npz:", {k: v.shape for k, v in out.items() if k.startswith("u8_")})


def golden_embeddings(E):
    """EmbeddingNet outputs of the reference for the MoCo ResNet-50 variants and the uber concatenations."""
    states = {n: restate.resnet50_state(VARIANT[n], s) for n, s in WEIGHT_SEEDS.items()}
    out = {"weight_seeds": np.array([WEIGHT_SEEDS[k] for k in ("moco_aug", "moco_aug_l4", "moco_aug_l3")])}
    frames64 = restate.structured_frames(4, 64, 64, 3, 21)
    frames224 = restate.structured_frames(2, 224, 224, 3, 22)
    obs2 = restate.structured_frames(3, 64, 64, 6, 23)  # ImageNav-style current||goal observation (n = 2)
    out.update(frames64=frames64, frames224=frames224, obs2=obs2)
    with tempfile.TemporaryDirectory() as d, refshim.chdir(d):
        refshim.write_checkpoints(d, states)
        for name in ("moco_aug", "moco_aug_l4", "moco_aug_l3", "moco_aug_uber_34", "moco_aug_uber_345"):
            net = E.EmbeddingNet(name, pretrained=True, train=False, disable_cuda=True)
            out[f"out_size_{name}"] = np.array(int(net.out_size))
            out[f"emb64_{name}"] = net(torch.from_numpy(frames64))
            if name in ("moco_aug", "moco_aug_uber_34"):
                out[f"emb224_{name}"] = net(torch.from_numpy(frames224))
            if name in ("moco_aug", "moco_aug_l3"):
                # the host loop of main_bc_1.py:128-137 on a 2-frame observation
                o = np.concatenate(np.split(obs2, 2, axis=3), axis=0)
                o = net(torch.from_numpy(o))
                out[f"emb_obs2_{name}"] = np.concatenate(np.split(o, 2, axis=0), axis=-1)
            # reference quirk D8: uber models have an empty state_dict
            out[f"n_state_keys_{name}"] = np.array(len(net.state_dict()))
            print(name, int(net.out_size), out[f"emb64_{name}"].shape, len(net.state_dict()))
    np.savez_compressed(os.path.join(GOLDEN, "embeddings.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    E = refshim.reference_embeddings()
    golden_transforms(E)
    golden_embeddings(E)


if __name__ == "__main__":
    main()
