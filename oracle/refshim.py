"""ORACLE (build container only) — import the UNMODIFIED reference from /root/reference.

The reference needs gym / detectron2 / timm / habitat at import time (src/embeddings.py:2-3,
src/vision_models/maskrcnn.py:2-20, src/vision_models/mae.py:20); none is installed. gym is not on the hot path and is replaced by an inert stub
module; detectron2's ResNet backbone classes (maskrcnn.py builds `maskrcnn_l3` from them) are served by the restatement
in oracle/restate_maskrcnn.py, the rest of detectron2 by inert classes; `timm.models.vision_transformer` (mae.py:20 needs PatchEmbed and Block
to build the MAE encoders) is served by the restatement of timm 0.5.4 in oracle/restate_mae.py. `np.float`, which
mae.py:58 still uses, left numpy in 1.24: it is restored as the alias of `float` it used to be.
Checkpoint-backed encoders (src/embeddings.py:151-236) read hard-coded relative paths: `write_checkpoints` writes synthetic checkpoints under those names into a scratch directory and the
constructors run from there, so no reference code is patched.
This file is never imported on the GPU box (/root/reference does not exist there).
"""
import contextlib
import os
import sys
import types

REFERENCE = "/root/reference"


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def install_stubs():
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class _W(object):
            def __init__(self, env=None):
                self.env = env

        gym.ObservationWrapper = gym.Wrapper = gym.Env = _W
        spaces = types.ModuleType("gym.spaces")
        box = types.ModuleType("gym.spaces.box")

        class Box(object):
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        box.Box = spaces.Box = Box
        gym.spaces = spaces
        spaces.box = box
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.box": box})
    if "detectron2" not in sys.modules:
        # the backbone classes mask_rcnn_model builds are served by the restatement of detectron2 in
        # oracle/restate_maskrcnn.py; everything else in the package (RPN, ROI heads, ...) is an inert class
        from oracle import restate_maskrcnn
        restate_maskrcnn.install_detectron2()
    for name in ("timm", "timm.models"):
        sys.modules.setdefault(name, _Anything(name))
    if "timm.models.vision_transformer" not in sys.modules:
        import numpy as np
        from oracle import restate_mae
        vt = types.ModuleType("timm.models.vision_transformer")
        vt.PatchEmbed, vt.Block = restate_mae.PatchEmbed, restate_mae.Block
        sys.modules["timm.models.vision_transformer"] = vt
        if not hasattr(np, "float"):
            np.float = float


def reference_embeddings():
    install_stubs()
    import src.embeddings as E
    return E


def reference_models():
    install_stubs()
    import src.models as M
    return M


CHECKPOINT_FILES = {  # src/embeddings.py:151-236
    "moco_aug": "moco_aug.pth.tar", "moco_aug_l4": "moco_aug_l4.pth", "moco_aug_l3": "moco_aug_l3.pth",
    "moco_croponly": "moco_croponly.pth", "moco_croponly_l4": "moco_croponly_l4.pth",
    "moco_croponly_l3": "moco_croponly_l3.pth",
}


def write_checkpoints(directory, states):
    """states: {embedding_name: torchvision-named state_dict}. Keys get the MoCo `module.encoder_q.` prefix
    (src/vision_models/moco.py:14-21)."""
    import torch
    for name, sd in states.items():
        ck = {"state_dict": {"module.encoder_q." + k: v.clone() for k, v in sd.items()}}
        torch.save(ck, os.path.join(directory, CHECKPOINT_FILES[name]))


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)
