"""MoCo-architecture ResNet-50 encoders — same constructors as the reference's src/vision_models/moco.py:6-113.

Each constructor loads `checkpoint['state_dict']`, keeps the `module.encoder_q.*` entries (minus `fc`), strips the
prefix and loads them non-strictly, asserting that nothing the architecture needs is missing — exactly the
reference's key handling (moco.py:14-24, 57-68, 100-111). The returned module holds the parameters under
torchvision's key names; the forward pass is the CUDA program built by pvr_habitat_b200.program.
"""
import contextlib
import os

import torch

from .resnet_params import ResNet50Params

_ALLOW_RANDOM_INIT = [False]


@contextlib.contextmanager
def allow_random_init():
    """Inside this context a missing checkpoint file leaves the network at its random initialisation
    (tests / benchmarks: there are no checkpoints offline). Outside it, a missing file raises like the reference."""
    _ALLOW_RANDOM_INIT.append(True)
    try:
        yield
    finally:
        _ALLOW_RANDOM_INIT.pop()


def random_init_allowed():
    return _ALLOW_RANDOM_INIT[-1]


def _load_encoder_q(model, checkpoint_path, allowed_unexpected):
    if not os.path.isfile(checkpoint_path) and _ALLOW_RANDOM_INIT[-1]:
        return model
    checkpoint = torch.load(checkpoint_path, map_location=torch.device('cpu'))
    state_dict = checkpoint['state_dict']
    for k in list(state_dict.keys()):
        # retain only encoder_q up to before the embedding layer (moco.py:14-21)
        if k.startswith('module.encoder_q') and not k.startswith('module.encoder_q.fc'):
            state_dict[k[len("module.encoder_q."):]] = state_dict[k]
        del state_dict[k]
    msg = model.load_state_dict(state_dict, strict=False)
    assert all(any(a in n for a in allowed_unexpected) for n in msg.unexpected_keys)
    assert len(msg.missing_keys) == 0
    return model


def moco_conv5(checkpoint_path):
    return _load_encoder_q(ResNet50Params('conv5'), checkpoint_path, ())


def moco_conv3_compressed(checkpoint_path):
    return _load_encoder_q(ResNet50Params('l3'), checkpoint_path, ('fc.', 'layer4.', 'layer3.2'))


def moco_conv4_compressed(checkpoint_path):
    return _load_encoder_q(ResNet50Params('l4'), checkpoint_path, ('fc.', 'layer4.2'))
