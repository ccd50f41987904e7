"""Host-side logic on CPU: the op program compiled from a state_dict computes the reference network (checked by
interpreting the exact fields the CUDA library consumes), names / sizes / state_dict keys mirror the reference."""
import numpy as np
import pytest
import torch

from oracle import restate
from program_emulator import emulate
from pvr_habitat_b200 import program as prg
from pvr_habitat_b200.embeddings import EmbeddingNet, UberModel, _get_embedding, resize_geometry
from pvr_habitat_b200.vision_models.moco import allow_random_init
from pvr_habitat_b200.vision_models.resnet_params import ResNet50Params


@pytest.mark.parametrize("variant,width", [("conv5", 2048), ("l4", 42 * 2 * 2), ("l3", 11 * 4 * 4)])
def test_program_equals_oracle_network(variant, width):
    sd = restate.resnet50_state(variant, 7)
    sd = {k: (v.to(torch.bfloat16).float() if k.endswith("conv1.weight") or "conv" in k or "downsample.0.weight" in k
              else v) for k, v in sd.items()}  # weights the program will round to bf16 anyway
    prog = prg.Program()
    s0 = prog.new_slot(64 * 32 * 32)
    prog.emb_width = prg.add_resnet50(prog, sd, variant, s0, 0, hw=64)
    assert prog.emb_width == width
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    x4 = torch.zeros(2, 64, 64, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    got = emulate(prog, prg.expand_stem_input(x4), round_bf16=False)
    ref = restate.resnet50_forward(sd, variant, x)
    # 3e-3: the four projection-shortcut blocks fold their BN scales into the packed weight rows, which are then
    # rounded to bf16 again (a packing mistake shows up at O(1))
    assert float((got - ref).abs().max()) <= 3e-3 * float(ref.abs().max())


def test_stem_packing_is_the_7x7_stride2_conv():
    w = torch.randn(64, 3, 7, 7)
    packed = prg.pack_stem_weight(w, 64).float().reshape(64, 8, 8, 4)  # (co, r, j, c)
    assert torch.all(packed[:, 7] == 0) and torch.all(packed[:, :, 7] == 0) and torch.all(packed[..., 3] == 0)
    assert torch.equal(packed[:, :7, :7, :3], w.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    # the expanded input layout: column e of output position q is input column 2q-3+e
    x4 = torch.arange(2 * 6 * 8 * 4, dtype=torch.float32).reshape(2, 6, 8, 4)
    ex = prg.expand_stem_input(x4).reshape(2, 6, 4, 8, 4)
    for q in range(4):
        for e in range(8):
            col = 2 * q - 3 + e
            want = x4[:, :, col] if 0 <= col < 8 else torch.zeros(2, 6, 4)
            assert torch.equal(ex[:, :, q, e], want)


def test_slot_planning_never_aliases_live_tensors():
    sd = restate.resnet50_state("conv5", 3)
    prog = prg.Program()
    s0 = prog.new_slot(224 * 112 * 32)
    prg.add_resnet50(prog, sd, "conv5", s0, 0)
    for op in prog.ops:
        if op["kind"] == 1:
            assert op["out_slot"] != op["in_slot"] and op["out_slot"] != 0
            if op["res_slot"] >= 0:
                assert op["res_slot"] != op["in_slot"]
            if op.get("in2_c", 0):
                assert op["in2_slot"] not in (op["out_slot"], op["in_slot"])
    # 53 convs - 4 projection shortcuts (fused into conv3 as a second K range) + maxpool + avgpool
    assert len(prog.ops) == 49 + 2
    assert sum(1 for op in prog.ops if op.get("in2_c", 0)) == 4
    assert len(prog.slot_elems) <= 8


def test_names_sizes_and_state_dict_keys():
    with allow_random_init():
        for name, size, nkeys in (("moco_aug", 2048, 318), ("moco_aug_l4", 2058, 337), ("moco_aug_l3", 2156, 277),
                                  ("moco_aug_uber_34", 4214, 0), ("moco_aug_uber_345", 6262, 0),
                                  ("moco_croponly_places_uber_45", 2058 + 2048, 0), ("resnet50", 2048, 318),
                                  ("resnet50_l3", 2156, 277)):
            net = EmbeddingNet(name, disable_cuda=True)
            assert net.out_size == size and tuple(net.in_shape) == (3, 224, 224)
            assert len(net.state_dict()) == nkeys  # uber: reference quirk D8 (plain list -> empty state_dict)
            assert all(k.startswith("embedding.") for k in net.state_dict())
    with pytest.raises(NotImplementedError, match="Requested model not available."):
        _get_embedding("no_such_model")


def test_state_dict_interchanges_with_reference_key_names():
    sd = restate.resnet50_state("l3", 5)
    m = ResNet50Params("l3")
    msg = m.load_state_dict(sd, strict=True)
    assert not msg.missing_keys and not msg.unexpected_keys


def test_missing_checkpoint_raises_like_reference():
    with pytest.raises(FileNotFoundError):
        EmbeddingNet("moco_aug", disable_cuda=True)


def test_cpu_forward_fails_loudly():
    with allow_random_init():
        net = EmbeddingNet("moco_aug", disable_cuda=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 64, 64, 3, dtype=torch.uint8))


def test_resize_geometry_matches_oracle():
    for hw in ((64, 64), (224, 224), (96, 128), (480, 640), (640, 480), (100, 75)):
        assert resize_geometry(*hw) == restate.resize_geometry(*hw)


def test_small_conv_program_equals_oracle(golden_dir):
    import os
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    sd = {k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w_")}
    sdb = {k: (v.to(torch.bfloat16).float() if k.endswith("weight") else v) for k, v in sd.items()}
    prog = prg.Program()
    s0 = prog.new_slot(64 * 64 * 4)
    prog.emb_width = prg.add_small_conv(prog, sd, s0, 0, hw=64)
    assert prog.emb_width == 32 * 2 * 2
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    x4 = torch.zeros(2, 64, 64, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    got = emulate(prog, x4, round_bf16=False)
    ref = restate.small_conv_forward(sdb, x)
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


def test_small_conv_oracle_matches_reference_golden(golden_dir):
    import os
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    sd = {k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w_")}
    for key in ("64", "224"):
        got = restate.small_conv_embedding(sd, gold["frames" + key])
        ref = gold["emb" + key]
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5 * float(np.abs(ref).max()))


def test_random_embedding_container_reproduces_reference_init(golden_dir):
    import os
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    torch.manual_seed(9)
    net = EmbeddingNet("random", pretrained=False, disable_cuda=True)
    assert net.out_size == int(gold["out_size"]) == 1568
    sd = net.state_dict()
    assert sorted(sd) == sorted("embedding." + k[2:] for k in gold.files if k.startswith("w_"))
    for k in gold.files:
        if k.startswith("w_"):  # orthogonal_ = LAPACK QR: last bits may differ between hosts
            np.testing.assert_allclose(sd["embedding." + k[2:]].numpy(), gold[k], atol=1e-5)


@pytest.mark.parametrize("name,n_ops", [("resnet18", 1 + 1 + 16 + 3 + 1), ("resnet34", 1 + 1 + 32 + 3 + 1)])
def test_resnet_basic_program_equals_oracle_network(name, n_ops):
    """resnet18 / resnet34 op programs (stem, max pool, BasicBlocks with 1x1 projection shortcuts, avg pool)."""
    from pvr_habitat_b200.vision_models.resnet_params import ResNetBasicParams
    sd = restate.resnet_basic_state(name, 4)
    sd = {k: (v.to(torch.bfloat16).float() if k.endswith(".weight") and v.dim() == 4 else v) for k, v in sd.items()}
    prog = prg.Program()
    s0 = prog.new_slot(64 * 32 * 32)
    prog.emb_width = prg.add_resnet_basic(prog, sd, restate.RESNET_BASIC_LAYERS[name], s0, 0, hw=64)
    assert prog.emb_width == 512 and len(prog.ops) == n_ops
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    x4 = torch.zeros(2, 64, 64, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    got = emulate(prog, prg.expand_stem_input(x4), round_bf16=False)
    ref = restate.resnet_basic_forward(sd, name, x)
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    # the parameter holder has torchvision's keys (checkpoints interchange)
    keys = set(ResNetBasicParams(name).state_dict().keys())
    assert keys == set(sd.keys())


# ------------------------------------------------------------------------------------------------ fp32 parity mode
def _frames4(n, hw, seed):
    x = torch.randn(n, 3, hw, hw, generator=torch.Generator().manual_seed(seed))
    x4 = torch.zeros(n, hw, hw, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    return x, x4


def _check_f32_program(prog):
    from pvr_habitat_b200 import _lib
    live = {}
    for op in prog.ops:
        assert op.get("flags", 0) & _lib.PVR_OP_FP32, "every op of an fp32 program carries PVR_OP_FP32"
        if op["kind"] == _lib.PVR_OP_CONV:
            assert op["k_pad"] == op["r"] * op["s"] * op["c_in"] and op["c_in"] % 4 == 0 and op["in_pitch"] % 4 == 0
            assert op["_weight"].dtype == torch.float32 and op["_weight"].shape == (op["c_out"], op["k_pad"])
            assert op["out_slot"] not in (op["in_slot"], op["res_slot"], 0)
            # slots count bf16 elements: a float32 tensor needs two per value
            assert prog.slot_elems[op["out_slot"]] >= 2 * op["h_out"] * op["w_out"] * op["out_pitch"]
            assert prog.slot_elems[op["in_slot"]] >= 2 * op["h_in"] * op["w_in"] * op["in_pitch"]
        live[op["out_slot"]] = op


@pytest.mark.parametrize("variant,width", [("conv5", 2048), ("l4", 42 * 2 * 2), ("l3", 11 * 4 * 4)])
def test_fp32_program_equals_oracle_network(variant, width):
    """The fp32 parity-mode program (PVR_OP_FP32 ops, dense float32 weights) is the oracle network to fp32 rounding."""
    sd = restate.resnet50_state(variant, 7)
    prog = prg.Program()
    s0 = prog.new_slot(64 * 64 * 4 * 2)
    prog.emb_width = prg.add_resnet50_f32(prog, sd, variant, s0, 0, hw=64)
    assert prog.emb_width == width
    _check_f32_program(prog)
    x, x4 = _frames4(2, 64, 0)
    got = emulate(prog, x4, round_bf16=False)
    ref = restate.resnet50_forward(sd, variant, x)
    assert float((got - ref).norm() / ref.norm()) <= 1e-5


@pytest.mark.parametrize("name", ["resnet18", "resnet34"])
def test_fp32_resnet_basic_program_equals_oracle_network(name):
    sd = restate.resnet_basic_state(name, 4)
    prog = prg.Program()
    s0 = prog.new_slot(64 * 64 * 4 * 2)
    prog.emb_width = prg.add_resnet_basic_f32(prog, sd, restate.RESNET_BASIC_LAYERS[name], s0, 0, hw=64)
    _check_f32_program(prog)
    x, x4 = _frames4(2, 64, 1)
    got = emulate(prog, x4, round_bf16=False)
    ref = restate.resnet_basic_forward(sd, name, x)
    assert float((got - ref).norm() / ref.norm()) <= 1e-5


def test_fp32_small_conv_program_equals_oracle(golden_dir):
    import os
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    sd = {k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w_")}
    prog = prg.Program()
    s0 = prog.new_slot(64 * 64 * 4 * 2)
    prog.emb_width = prg.add_small_conv_f32(prog, sd, s0, 0, hw=64)
    _check_f32_program(prog)
    x, x4 = _frames4(2, 64, 1)
    got = emulate(prog, x4, round_bf16=False)
    ref = restate.small_conv_forward(sd, x)
    assert float((got - ref).norm() / ref.norm()) <= 1e-5


def test_set_precision_validates_and_invalidates():
    with allow_random_init():
        net = EmbeddingNet("moco_aug", disable_cuda=True)
    assert net.precision == "bf16"
    net._encoder = object()
    assert net.set_precision("fp32") is net and net._encoder is None and net.precision == "fp32"
    with pytest.raises(ValueError):
        net.set_precision("fp16")
    with allow_random_init():
        clip = EmbeddingNet("clip_vit", disable_cuda=True)
    assert clip.set_precision("fp32").precision == "fp32"  # the ViT encoders have the fp32 mode too (csrc/vit_f32.cu)


# ------------------------------------------------------------------------------------------------ transforms dispatch
def test_transforms_dispatch_per_interpolation(monkeypatch):
    """Which C entry point / format flag `Transforms.run` uses: bilinear (default), bicubic (MAE: flag on out_fmt),
    antialiased bicubic (CLIP: own entry point, except for the identity resize, which stays on the plain kernel)."""
    import contextlib
    from pvr_habitat_b200 import _lib
    from pvr_habitat_b200.embeddings import CLIP_MEAN, CLIP_STD, Transforms
    calls = []

    class FakeLib:
        def pvr_preprocess_u8(self, *a):
            calls.append(("plain",) + a)
            return 0

        def pvr_preprocess_u8_aa(self, *a):
            calls.append(("aa",) + a)
            return 0

    monkeypatch.setattr(_lib, "lib", lambda: FakeLib())
    monkeypatch.setattr(_lib, "current_stream_ptr", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    obs64 = torch.zeros(2, 64, 64, 6, dtype=torch.uint8)
    obs224 = torch.zeros(1, 224, 224, 3, dtype=torch.uint8)
    Transforms().run(obs64, 2, 1234, _lib.PVR_FMT_STEM_BF16, True)
    Transforms(interpolation="bicubic").run(obs64, 2, 1234, _lib.PVR_FMT_NHWC4_BF16, True)
    clip = Transforms(CLIP_MEAN, CLIP_STD, size=224, crop=224, interpolation="bicubic_aa")
    clip.run(obs64, 2, 1234, _lib.PVR_FMT_NHWC4_BF16, True)
    clip.run(obs224, 1, 1234, _lib.PVR_FMT_NHWC4_BF16, True)
    kinds = [c[0] for c in calls]
    assert kinds == ["plain", "plain", "aa", "plain"]
    # (in, N, H, W, n_frames, rh, rw, top, left, crop, mean, std, out, fmt, sample_major, stream)
    assert calls[0][2:11] == (2, 64, 64, 2, 256, 256, 16, 16, 224) and calls[0][14] == _lib.PVR_FMT_STEM_BF16
    assert calls[1][14] == _lib.PVR_FMT_NHWC4_BF16 | _lib.PVR_RESIZE_BICUBIC
    assert calls[2][2:11] == (2, 64, 64, 2, 224, 224, 0, 0, 224) and calls[2][14] == _lib.PVR_FMT_NHWC4_BF16
    assert calls[3][2:11] == (1, 224, 224, 1, 224, 224, 0, 0, 224) and calls[3][14] == _lib.PVR_FMT_NHWC4_BF16
    assert [round(v, 4) for v in calls[2][11]] == [round(v, 4) for v in CLIP_MEAN]
    # CLIP nets take any frame size (antialiased bicubic kernel, verified on B200 in round 2)
    with allow_random_init():
        net = EmbeddingNet("clip_vit", disable_cuda=True)
    assert net.transforms.interpolation == "bicubic_aa" and not net.transforms.identity_resize_only


# ------------------------------------------------------------------------------------------------ name coverage
REFERENCE_NAMES = """clip_rn50 clip_vit demy mae_base mae_huge mae_large maskrcnn_l3 moco_aug moco_aug_habitat moco_aug_l3
moco_aug_l4 moco_aug_mujoco moco_aug_places moco_aug_places_l3 moco_aug_places_l4 moco_aug_places_uber_34
moco_aug_places_uber_345 moco_aug_places_uber_35 moco_aug_places_uber_45 moco_aug_uber moco_aug_uber_34 moco_aug_uber_345
moco_aug_uber_35 moco_aug_uber_45 moco_coloronly moco_croponly moco_croponly_habitat moco_croponly_l3 moco_croponly_l4
moco_croponly_mujoco moco_croponly_places moco_croponly_places_l3 moco_croponly_places_l4 moco_croponly_places_uber_34
moco_croponly_places_uber_345 moco_croponly_places_uber_35 moco_croponly_places_uber_45 moco_croponly_uber
moco_croponly_uber_34 moco_croponly_uber_345 moco_croponly_uber_35 moco_croponly_uber_45 random resnet18 resnet34 resnet50
resnet50_l3 resnet50_l4 resnet50_places resnet50_places_l3 resnet50_places_l4 true_state""".split()
NOT_BUILT = set()  # every name of the reference builds


def test_every_reference_embedding_name_is_accounted_for():
    """The 52 names `_get_embedding` accepts in the reference (every `embedding_name == '...'` of src/embeddings.py:
    88-318): all 52 build here with the reference's output width; an unknown name raises NotImplementedError like an unknown name."""
    widths = {"conv5": 2048, "l4": 2058, "l3": 2156}
    assert len(REFERENCE_NAMES) == 52
    with allow_random_init():
        for name in REFERENCE_NAMES:
            if name in NOT_BUILT:
                with pytest.raises(NotImplementedError):
                    _get_embedding(name)
                continue
            model, transforms = _get_embedding(name)
            if name == "true_state":
                continue
            if "_uber_" in name:
                want = sum(widths[{"3": "l3", "4": "l4", "5": "conv5"}[t]] for t in name.split("_uber_")[1])
            elif name.endswith(("_l3", "_l4")):  # maskrcnn_l3 included: (11, 14, 14)
                want = widths[name[-2:]]
            else:
                want = {"random": 1568, "resnet18": 512, "resnet34": 512, "clip_vit": 512, "clip_rn50": 1024, "mae_base": 768,
                        "mae_large": 1024, "mae_huge": 1280}.get(name, 2048)
            assert int(model.out_size) == want, name
            assert not model.training and all(not p.requires_grad for p in model.parameters()), name
    with pytest.raises(NotImplementedError):
        _get_embedding("no_such_model")


def test_embedding_wrapper_without_gym():
    """EmbeddingWrapper (src/embeddings.py:409-444) works on a duck-typed environment when gym is not installed: the
    observation space becomes (n_frames * out_size,), reset / step route observations through the encoder."""
    import types
    from pvr_habitat_b200.embeddings import EmbeddingWrapper

    class Enc:
        out_size = 5
        calls = []

        def embed(self, obs, n_frames):
            Enc.calls.append((tuple(obs.shape), obs.dtype, n_frames))
            return torch.arange(n_frames * 5, dtype=torch.float32).reshape(1, -1)

    class Env:
        observation_space = types.SimpleNamespace(shape=(64, 64, 6))
        action_space = types.SimpleNamespace(n=3)

        def reset(self):
            return np.zeros((64, 64, 6), np.uint8)

        def step(self, a):
            return np.ones((64, 64, 6), np.uint8), 1.0, False, {}

    w = EmbeddingWrapper(Env(), Enc())
    assert w.n_frames == 2 and tuple(w.observation_space.shape) == (10,)
    o = w.reset()
    assert o.shape == (10,) and o.dtype == np.float32 and Enc.calls[-1] == ((1, 64, 64, 6), torch.uint8, 2)
    o, r, d, info = w.step(0)
    assert o.shape == (10,) and r == 1.0 and d is False and w.action_space.n == 3


# ------------------------------------------------------------------------------------------------ encoders of §8(f)-4
def test_maskrcnn_program_equals_oracle_backbone():
    """detectron2-named weights -> program_state_dict -> add_resnet50(variant='l3', stride_in_1x1=True), emulated on the
    CPU, against the restated detectron2 backbone (stride on the first 1x1, 1x1 shortcut of the compression block as the
    centre tap of a 3x3 kernel)."""
    from oracle import restate_maskrcnn as rm
    from pvr_habitat_b200.vision_models.maskrcnn import MaskRCNNBackboneParams
    sd = rm.maskrcnn_state(5)
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() == 4 else v) for k, v in sd.items()}
    prog = prg.Program()
    s0 = prog.new_slot(64 * 32 * 32)
    prog.emb_width = prg.add_resnet50(prog, MaskRCNNBackboneParams.program_state_dict(sd), "l3", s0, 0, hw=64,
                                      stride_in_1x1=True)
    assert prog.emb_width == 11 * 4 * 4
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1)) * 40
    x4 = torch.zeros(2, 64, 64, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    got = emulate(prog, prg.expand_stem_input(x4), round_bf16=False)
    net = rm.build_backbone()
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        ref = net(x)["res4"].reshape(2, -1)
    assert float((got - ref).abs().max()) <= 3e-3 * float(ref.abs().max())


def test_clip_resnet_program_equals_oracle_trunk():
    """add_clip_resnet (3-conv stem, PVR_OP_AVGPOOL2, pooled projection blocks as one GEMM over [t2 | x]) emulated on the
    CPU against the restated ModifiedResNet up to its attention pool; the slot planner must keep the feature map the
    runner reads alive."""
    from oracle import restate_clip_rn as rc
    sd = {k[len("visual."):]: v for k, v in rc.clip_rn50_state(3).items() if k.startswith("visual.")}
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() == 4 else v) for k, v in sd.items()}
    prog = prg.Program()
    s0 = prog.new_slot(64 * 64 * 4)
    feat, chw = prg.add_clip_resnet(prog, sd, s0, hw=64)
    prog.emb_width = 1
    assert chw == (2048, 2, 2)
    kinds = [op["kind"] for op in prog.ops]
    assert kinds.count(_lib_kinds()["avgpool2"]) == 1 + 2 * 3  # stem + (t2, x) of layer2.0 / layer3.0 / layer4.0
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    x4 = torch.zeros(2, 64, 64, 4)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    _, slots = emulate(prog, x4, round_bf16=False, return_slots=True)
    got = slots[feat][:, :2 * 2 * 2048].reshape(2, 2, 2, 2048).permute(0, 3, 1, 2)
    m = rc.ModifiedResNet().eval()
    m.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    with torch.no_grad():
        t = x
        for conv, bn in ((m.conv1, m.bn1), (m.conv2, m.bn2), (m.conv3, m.bn3)):
            t = torch.relu(bn(conv(t)))
        ref = m.layer4(m.layer3(m.layer2(m.layer1(m.avgpool(t)))))
    assert float((got - ref).abs().max()) <= 3e-3 * float(ref.abs().max())
    # no conv writes into a slot it (or its residual / second input) still reads
    for op in prog.ops:
        if op["kind"] == 1:
            assert op["out_slot"] not in (op["in_slot"], 0)
            if op.get("in2_c", 0):
                assert op["in2_slot"] not in (op["out_slot"], op["in_slot"])


def _lib_kinds():
    from pvr_habitat_b200 import _lib
    return {"avgpool2": _lib.PVR_OP_AVGPOOL2}
