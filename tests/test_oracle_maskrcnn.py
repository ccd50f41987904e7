"""maskrcnn_l3 (src/embeddings.py:283-295, src/vision_models/maskrcnn.py): the oracle restatement against the goldens
that oracle/make_golden.py wrote by running the UNMODIFIED reference EmbeddingNet('maskrcnn_l3') (its transforms, stage
surgery and `['res4']` tap) on top of the detectron2 restatement of oracle/restate_maskrcnn.py, and the host side of the
drop-in (key names, checkpoint filter). detectron2 itself is not installed: parity with detectron2's own code is
unpinned. CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import restate_maskrcnn as rm
from pvr_habitat_b200.embeddings import EmbeddingNet
from pvr_habitat_b200.vision_models import maskrcnn
from pvr_habitat_b200.vision_models.moco import allow_random_init

CASES = ["structured_64", "structured_224", "small_40x48", "adversarial_64"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "maskrcnn.npz"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_transforms_bit_exact_vs_reference(gold, case):
    t = rm.maskrcnn_transforms(gold["in_" + case])
    assert np.array_equal(t[:, :, :12], gold["t_top_" + case])
    assert np.array_equal(t[:, :, ::7, ::5], gold["t_sub_" + case])


def test_row_permutation_reaches_the_crop_only_for_short_frames(gold):
    """`x[:,:,[0,1,2]] = x[:,:,[2,1,0]]` permutes rows, not channels: visible after CenterCrop only below 54 rows."""
    for case, visible in (("small_40x48", True), ("structured_64", False)):
        frames = gold["in_" + case]
        plain = frames.copy()
        plain[:, [0, 2]] = plain[:, [2, 0]]  # undo: the oracle applied to pre-swapped frames = no permutation
        assert (not np.array_equal(rm.maskrcnn_transforms(plain), rm.maskrcnn_transforms(frames))) == visible


@pytest.mark.parametrize("case", CASES)
def test_oracle_embedding_matches_reference(gold, case):
    sd = rm.maskrcnn_state(int(gold["seed"]))
    got = rm.embedding_forward(sd, gold["in_" + case])
    ref = gold["emb_" + case]
    assert got.shape == ref.shape == (len(gold["in_" + case]), 2156)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6


def test_container_keys_and_checkpoint_filter(gold, tmp_path):
    with allow_random_init():
        net = EmbeddingNet("maskrcnn_l3", disable_cuda=True)
    assert net.out_size == int(gold["out_size"]) == 2156 and tuple(net.in_shape) == (3, 224, 224) and not net.training
    assert sorted(net.state_dict().keys()) == list(gold["state_keys"])
    assert net.transforms.interpolation == "bilinear_float_rows02" and net.transforms.mean == rm.PIXEL_MEAN
    with pytest.raises(FileNotFoundError):
        EmbeddingNet("maskrcnn_l3", disable_cuda=True)
    # a reference checkpoint holds the whole GeneralizedRCNN: heads and the emptied res4.7 block are dropped
    sd = rm.maskrcnn_state(3)
    full = {"backbone." + k: v for k, v in sd.items()}
    full.update({"backbone.res4.7.conv1.weight": torch.zeros(1024, 11, 3, 3), "roi_heads.box_predictor.cls_score.weight":
                 torch.zeros(81, 2048), "proposal_generator.rpn_head.conv.weight": torch.zeros(1024, 1024, 3, 3)})
    path = str(tmp_path / "maskrcnn_l3.pth")
    torch.save({"model": full}, path)
    m = maskrcnn.mask_rcnn_model(path)
    got = m.state_dict()
    assert all(torch.equal(got[k], v) for k, v in sd.items()) and len(got) == len(sd)
    # translation into the names program.add_resnet50(variant='l3', stride_in_1x1=True) reads
    psd = m.program_state_dict(got)
    assert torch.equal(psd["layer3.0.5.conv3.weight"], sd["res4.5.conv3.weight"])
    assert torch.equal(psd["layer2.0.downsample.1.running_var"], sd["res3.0.shortcut.norm.running_var"])
    w3 = psd["layer3.1.downsample.0.weight"]
    assert w3.shape == (11, 1024, 3, 3) and torch.equal(w3[:, :, 1, 1], sd["res4.6.shortcut.weight"][:, :, 0, 0])
    assert float(w3.abs().sum()) == float(sd["res4.6.shortcut.weight"].abs().sum())
    assert not psd["layer3.1.downsample.0.bias"].any()
