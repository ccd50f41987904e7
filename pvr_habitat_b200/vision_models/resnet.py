"""Supervised ResNet-50 encoders — same constructors as the reference's src/vision_models/resnet.py:6-104
(checkpoints whose keys carry a `module.` DataParallel prefix; compressed variants share moco.py's architecture)."""
import os

import torch

from .moco import _ALLOW_RANDOM_INIT
from .resnet_params import ResNet50Params


def _load_module_prefixed(model, checkpoint_path, allowed_unexpected, check_missing):
    if not os.path.isfile(checkpoint_path) and _ALLOW_RANDOM_INIT[-1]:
        return model
    checkpoint = torch.load(checkpoint_path, map_location=torch.device('cpu'))
    state_dict = checkpoint['state_dict']
    for k in list(state_dict.keys()):
        if k.startswith('module.'):  # resnet.py:35-39
            state_dict[k[len('module.'):]] = state_dict[k]
        del state_dict[k]
    msg = model.load_state_dict(state_dict, strict=False)
    if allowed_unexpected is not None:
        assert all(any(a in n for a in allowed_unexpected) for n in msg.unexpected_keys)
    if check_missing:
        assert len(msg.missing_keys) == 0
    return model


def resnet_conv3_compressed(checkpoint_path):
    return _load_module_prefixed(ResNet50Params('l3'), checkpoint_path, ('fc.', 'layer4.', 'layer3.2'), False)


def resnet_conv4_compressed(checkpoint_path):
    return _load_module_prefixed(ResNet50Params('l4'), checkpoint_path, ('fc.', 'layer4.2'), False)


def resnet_conv5(checkpoint_path):
    return _load_module_prefixed(ResNet50Params('conv5'), checkpoint_path, None, True)
