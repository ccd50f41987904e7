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


def golden_small_conv(E):
    """The reference's 'random' PVR (5-layer conv, src/embeddings.py:90-106); its weights are small enough to store."""
    torch.manual_seed(9)
    net = E.EmbeddingNet('random', pretrained=False, train=False, disable_cuda=True)
    frames64 = restate.structured_frames(4, 64, 64, 3, 51)
    frames224 = restate.structured_frames(2, 224, 224, 3, 52)
    out = {"frames64": frames64, "frames224": frames224, "out_size": np.array(int(net.out_size)),
           "emb64": net(torch.from_numpy(frames64)), "emb224": net(torch.from_numpy(frames224))}
    for k, v in net.embedding.state_dict().items():
        out["w_" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "small_conv.npz"), **out)
    print("small_conv.npz", out["emb64"].shape, int(net.out_size))


RESNET_BASIC_SEEDS = {"resnet18": 61, "resnet34": 62}


def golden_resnet_basic(E):
    """The reference's `resnet18` / `resnet34` encoders (torchvision nets with fc = Identity, src/embeddings.py:112-117)
    on synthetic weights from `restate.resnet_basic_state(name, seed)` (loaded with strict=True)."""
    frames64 = restate.structured_frames(4, 64, 64, 3, 71)
    frames224 = restate.structured_frames(2, 224, 224, 3, 72)
    out = {"frames64": frames64, "frames224": frames224}
    for name, seed in RESNET_BASIC_SEEDS.items():
        net = E.EmbeddingNet(name, pretrained=False, train=False, disable_cuda=True)
        net.embedding.load_state_dict(restate.resnet_basic_state(name, seed), strict=True)
        out[f"seed_{name}"] = np.array(seed)
        out[f"out_size_{name}"] = np.array(int(net.out_size))
        out[f"emb64_{name}"] = net(torch.from_numpy(frames64))
        out[f"emb224_{name}"] = net(torch.from_numpy(frames224))
        out[f"n_state_keys_{name}"] = np.array(len(net.state_dict()))
        print(name, int(net.out_size), out[f"emb64_{name}"].shape, len(net.state_dict()))
    np.savez_compressed(os.path.join(GOLDEN, "resnet_basic.npz"), **out)


def golden_policy():
    """Reference PolicyNet (src/models.py) forward/backward, and the per-step training trace of the UNMODIFIED
    main_bc_2.run() on a synthetic embedded-observation pickle (fake env / test modules)."""
    import pickle
    import random
    import types
    from oracle import restate_policy as rp
    M = refshim.reference_models()
    out = {}
    # ---- (a) one forward/backward of the reference module
    T, B, D, A = 6, 5, 64, 3
    torch.manual_seed(7)
    net = M.PolicyNet((D,), A, batch_norm=True)
    net.train()
    rng = np.random.default_rng(3)
    obs = rng.standard_normal((T, B, D)).astype(np.float32)
    done = rng.random((T, B)) < 0.25
    act = rng.integers(0, A, (T, B))
    state = tuple(torch.from_numpy(rng.standard_normal((2, B, 1024)).astype(np.float32)) * 0.1 for _ in range(2))
    o, st = net(dict(obs=torch.from_numpy(obs), done=torch.from_numpy(done)), state)
    loss = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(torch.flatten(o["policy_logits"], 0, 1), -1),
                                        torch.flatten(torch.from_numpy(act), 0, 1).long())
    loss.backward()
    out.update(fb_obs=obs, fb_done=done, fb_act=act, fb_h0=state[0].numpy(), fb_c0=state[1].numpy(),
               fb_logits=o["policy_logits"].detach().numpy(), fb_baseline=o["baseline"].detach().numpy(),
               fb_hn=st[0].detach().numpy(), fb_cn=st[1].detach().numpy(), fb_loss=np.float32(loss.item()))
    names, norms, none = [], [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(float(p.grad.norm()) if p.grad is not None else -1.0)
        if p.grad is None:
            none.append(k)
    out.update(fb_param_names=np.array(names), fb_grad_norms=np.array(norms, dtype=np.float64),
               fb_param_sums=np.array([float(p.double().sum()) for p in net.parameters()]),
               fb_grad_bias_ih_l1=net.core.bias_ih_l1.grad.numpy(), fb_grad_policy_w=net.policy.weight.grad.numpy(),
               fb_grad_bn_w=net.fc[0].weight.grad.numpy(), fb_running_var=net.fc[0].running_var.numpy())
    print("policy fwd/bwd: loss", loss.item(), "params without grad:", none)

    # ---- (a2) PolicyNetWithConv (src/models.py:96-197) forward/backward on 2-frame 64x64 observations
    T, B, A = 3, 4, 3
    torch.manual_seed(13)
    netc = M.PolicyNetWithConv((64, 64, 6), A, batch_norm=True)
    netc.train()
    obs_c = restate.structured_frames(T * B, 64, 64, 6, 61).reshape(T, B, 64, 64, 6)
    done_c = np.random.default_rng(4).random((T, B)) < 0.25
    act_c = np.random.default_rng(5).integers(0, A, (T, B))
    oc, stc = netc(dict(obs=torch.from_numpy(obs_c), done=torch.from_numpy(done_c)), netc.initial_state(B))
    loss_c = torch.nn.functional.nll_loss(
        torch.nn.functional.log_softmax(torch.flatten(oc["policy_logits"], 0, 1), -1),
        torch.flatten(torch.from_numpy(act_c), 0, 1).long())
    loss_c.backward()
    out.update(cv_obs=obs_c, cv_done=done_c, cv_act=act_c, cv_logits=oc["policy_logits"].detach().numpy(),
               cv_loss=np.float32(loss_c.item()),
               cv_param_names=np.array([k for k, _ in netc.named_parameters()]),
               cv_param_sums=np.array([float(p.double().sum()) for p in netc.parameters()]),
               cv_grad_norms=np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                       for p in netc.parameters()], dtype=np.float64),
               cv_grad_conv0_w=netc.feat_extract[0].weight.grad.numpy(),
               cv_grad_conv4_w=netc.feat_extract[8].weight.grad.numpy(),
               cv_grad_conv2_b=netc.feat_extract[4].bias.grad.numpy())
    print("PolicyNetWithConv fwd/bwd: loss", loss_c.item())

    # ---- (b) main_bc_2.run() unmodified
    n, D, T, B, steps = 512, 128, 8, 4, 12
    obs, action, done, reward = rp.synthetic_bc_data(n, D, 3, 11)
    fake_env = types.ModuleType("src.env_utils")

    class _Space:
        def __init__(self, shape=None, n=None):
            self.shape, self.n = shape, n

    class _Env:
        def __init__(self):
            self.gym_env = types.SimpleNamespace(observation_space=_Space(shape=(D,)), action_space=_Space(n=3))

        def close(self):
            pass

    fake_env.make_environment = lambda flags, embedding_model, actor_id=1: _Env()
    fake_test = types.ModuleType("src.test_model")
    fake_test.test = lambda model, env, stat_keys, n_episodes: {k: [0.0] for k in stat_keys}
    sys.modules["src.env_utils"], sys.modules["src.test_model"] = fake_env, fake_test
    refshim.install_stubs()
    import src.embeddings as E
    E.EmbeddingNet = lambda *a, **k: torch.nn.Identity()
    import main_bc_2
    main_bc_2.EmbeddingNet = E.EmbeddingNet
    from src.arguments import parser
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "fakeenv_fakeemb.pickle"), "wb") as f:
            pickle.dump(dict(obs=obs, action=action, reward=reward, done=done, true_state=np.zeros((n, 12))), f)
        flags = parser.parse_args([
            "--env", "fakeenv", "--to_env", "fakeenv", "--embedding_name", "fakeemb", "--data_path", d,
            "--save_path", os.path.join(d, "out"), "--batch_size", str(B), "--unroll_length", str(T),
            "--max_frames", str(steps * T * B), "--eval_frequency", "1", "--batch_norm", "--disable_cuda",
            "--disable_save", "--run_id", "5", "--n_episodes_test", "1"])
        stats = {}
        orig_dump = pickle.dump
        main_bc_2.pickle.dump = lambda obj, fh, protocol=None: stats.update(obj)  # unused (disable_save)
        # capture the stats dict: run() keeps it local, so re-enable saving into the temp dir instead
        flags.disable_save = False
        main_bc_2.pickle.dump = orig_dump
        main_bc_2.run(flags)
        st = pickle.load(open(os.path.join(d, "out", "fakeenv_emfakeemb_s5_fakeenv.pickle"), "rb"))["fakeenv"]
    out.update(bc_obs=obs, bc_action=action, bc_done=done, bc_T=np.array(T), bc_B=np.array(B),
               bc_steps=np.array(steps), bc_seed=np.array(5), bc_max_frames=np.array(steps * T * B),
               bc_loss=np.array(st["training_loss"][1:], dtype=np.float64),
               bc_grad_norm=np.array(st["gradient_norm"][1:], dtype=np.float64),
               bc_frames=np.array(st["frames"][1:]))
    print("main_bc_2.run trace:", np.round(out["bc_loss"], 4))
    np.savez_compressed(os.path.join(GOLDEN, "policy.npz"), **out)


def golden_mae(E):
    """EmbeddingNet('mae_base' / 'mae_large') of the reference (src/embeddings.py:81,137-144,377-379) on synthetic
    checkpoints written under the hard-coded file names, plus the uint8 output of its bicubic Resize + CenterCrop.

    torchvision 0.26 would antialias the bicubic resize (default antialias=True, a different filter: A = -0.5); the
    reference pins torchvision 0.10 whose tensor path has no antialiasing, so the constructed Resize module gets
    `antialias = False` — an attribute of the torchvision object, no reference code is touched."""
    from oracle import restate_mae
    out = {}
    cases = {
        "structured_64": restate.structured_frames(2, 64, 64, 3, 41),
        "structured_224": restate.structured_frames(2, 224, 224, 3, 42),
        "structured_96x128": restate.structured_frames(1, 96, 128, 3, 43),
        "noise_224": np.random.default_rng(44).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8),
        "adversarial_64": restate.adversarial_frames(64, 64),
    }
    frames224 = restate.structured_frames(2, 224, 224, 3, 45)
    frames64 = restate.structured_frames(3, 64, 64, 3, 46)
    out.update(frames224=frames224, frames64=frames64)
    seeds = {"mae_base": 201, "mae_large": 202}
    for name, seed in seeds.items():
        sd = restate_mae.mae_state(name, seed)
        with tempfile.TemporaryDirectory() as d, refshim.chdir(d):
            torch.save({"model": sd}, os.path.join(d, restate_mae.CHECKPOINTS[name]))
            torch.manual_seed(7)  # random_masking draws torch.rand even at mask_ratio 0
            net = E.EmbeddingNet(name, pretrained=True, train=False, disable_cuda=True)
        net.transforms[0].antialias = False
        out[f"seed_{name}"] = np.array(seed)
        out[f"out_size_{name}"] = np.array(int(net.out_size))
        out[f"emb224_{name}"] = net(torch.from_numpy(frames224))
        out[f"emb64_{name}"] = net(torch.from_numpy(frames64))
        if name == "mae_base":
            resize_crop = torch.nn.Sequential(net.transforms[0], net.transforms[1])
            for cname, frames in cases.items():
                x = torch.from_numpy(frames).permute(0, 3, 1, 2).contiguous()
                u = resize_crop(x)
                assert u.dtype == torch.uint8
                out["in_" + cname] = frames
                out["u8_" + cname] = u.numpy()
        print(name, out[f"out_size_{name}"], np.abs(out[f"emb224_{name}"]).mean())
    # the reference's random initialisation (mae.py:117-146) under a fixed torch seed: slices of a few tensors pin the
    # order of the random draws that the drop-in parameter container has to reproduce
    from src.vision_models.mae import mae_vit_base_patch16
    torch.manual_seed(9)
    ref = mae_vit_base_patch16().state_dict()
    out["init_keys"] = np.array(sorted(ref.keys()))
    for k in ("cls_token", "mask_token", "patch_embed.proj.weight", "blocks.0.attn.qkv.weight",
              "blocks.11.mlp.fc2.weight", "decoder_blocks.7.mlp.fc1.weight", "decoder_pred.weight", "pos_embed",
              "decoder_pos_embed"):
        out["init_" + k] = ref[k].reshape(-1)[:: max(1, ref[k].numel() // 64)][:64].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "mae.npz"), **out)


def golden_mae_huge(E):
    """EmbeddingNet('mae_huge') of the reference (src/embeddings.py:145-148: mae_vit_huge_patch14, 257 tokens, 16 heads
    of 80) on a synthetic checkpoint, in its own file so that mae.npz stays as generated. One frame of each of the
    frame sets of golden_mae (632 M parameters: 0.33 TFLOP per frame on the CPU)."""
    from oracle import restate_mae
    name, seed = "mae_huge", 203
    frames224 = restate.structured_frames(2, 224, 224, 3, 45)[:1]
    frames64 = restate.structured_frames(3, 64, 64, 3, 46)[:1]
    sd = restate_mae.mae_state(name, seed)
    with tempfile.TemporaryDirectory() as d, refshim.chdir(d):
        torch.save({"model": sd}, os.path.join(d, restate_mae.CHECKPOINTS[name]))
        torch.manual_seed(7)
        net = E.EmbeddingNet(name, pretrained=True, train=False, disable_cuda=True)
    net.transforms[0].antialias = False  # torchvision 0.10 semantics, see golden_mae
    out = {"frames224": frames224, "frames64": frames64, f"seed_{name}": np.array(seed),
           f"out_size_{name}": np.array(int(net.out_size)),
           f"emb224_{name}": np.atleast_2d(net(torch.from_numpy(frames224))),
           f"emb64_{name}": np.atleast_2d(net(torch.from_numpy(frames64)))}
    print(name, out[f"out_size_{name}"], np.abs(out[f"emb224_{name}"]).mean())
    np.savez_compressed(os.path.join(GOLDEN, "mae_huge.npz"), **out)


def golden_maskrcnn(E):
    """EmbeddingNet('maskrcnn_l3') of the reference (src/embeddings.py:283-295, :380-383; src/vision_models/maskrcnn.py)
    on a synthetic checkpoint: its transforms (torchvision) and model surgery run unmodified; detectron2's backbone
    classes come from the restatement in oracle/restate_maskrcnn.py (detectron2 is not installed).
    torchvision 0.26 would route the float Resize through the antialiased kernel (same weights when up-scaling, another
    summation order); the reference pins torchvision 0.10: `antialias = False` on the constructed Resize object."""
    from oracle import restate_maskrcnn
    seed = 301
    sd = restate_maskrcnn.maskrcnn_state(seed)
    full = {"backbone." + k: v for k, v in sd.items()}
    g = torch.Generator().manual_seed(seed)
    for k, v in restate_maskrcnn.BasicBlock(11, 1024).state_dict().items():  # res4[7]: loaded strictly, then dropped
        full["backbone.res4.7." + k] = torch.randn(v.shape, generator=g) if v.dtype.is_floating_point else v
    with tempfile.TemporaryDirectory() as d, refshim.chdir(d):
        torch.save({"model": full}, os.path.join(d, "maskrcnn_l3.pth"))
        net = E.EmbeddingNet("maskrcnn_l3", pretrained=True, train=False, disable_cuda=True)
    net.transforms[1].antialias = False
    cases = {
        "structured_64": restate.structured_frames(2, 64, 64, 3, 71),
        "structured_224": restate.structured_frames(1, 224, 224, 3, 72),
        "small_40x48": restate.structured_frames(1, 40, 48, 3, 73),   # < 54 rows: the permuted rows 0 / 2 reach the crop
        "adversarial_64": restate.adversarial_frames(64, 64)[:2],
    }
    out = {"seed": np.array(seed), "out_size": np.array(int(net.out_size)),
           "state_keys": np.array(sorted(net.state_dict().keys()))}
    for name, frames in cases.items():
        x = torch.from_numpy(frames).permute(0, 3, 1, 2).contiguous()
        t = net.transforms(x.clone()).numpy()
        out["in_" + name] = frames
        out["t_top_" + name] = t[:, :, :12]          # the rows the reference's row permutation can reach
        out["t_sub_" + name] = t[:, :, ::7, ::5]
        out["emb_" + name] = np.atleast_2d(net(torch.from_numpy(frames)))
        print(name, t.shape, float(np.abs(out["emb_" + name]).mean()))
    np.savez_compressed(os.path.join(GOLDEN, "maskrcnn.npz"), **out)


def golden_clip_rn50(E):
    """EmbeddingNet('clip_rn50') of the reference (src/embeddings.py:305-314, 375-376): its transforms, `encode_image`
    dispatch and output handling run unmodified; `clip.load("RN50")` is served by the restatement of openai/CLIP's
    ModifiedResNet in oracle/restate_clip_rn.py (the `clip` package is not installed)."""
    import types
    from oracle import restate_clip_rn
    seed = 401
    sd = restate_clip_rn.clip_rn50_state(seed)

    def load(name, device="cpu"):
        assert name == "RN50"
        m = restate_clip_rn.FakeClip()
        m.load_state_dict(sd, strict=True)
        return m.to(device), None

    fake = types.ModuleType("clip")
    fake.load = load
    old = getattr(E, "clip", None)
    E.clip = fake
    try:
        net = E.EmbeddingNet("clip_rn50", pretrained=True, train=False, disable_cuda=True)
    finally:
        if old is not None:
            E.clip = old
    cases = {"structured_64": restate.structured_frames(2, 64, 64, 3, 81),
             "structured_224": restate.structured_frames(1, 224, 224, 3, 82),
             "structured_96x128": restate.structured_frames(1, 96, 128, 3, 83)}
    out = {"seed": np.array(seed), "out_size": np.array(int(net.out_size)),
           "visual_keys": np.array(sorted(k for k in net.state_dict() if k.startswith("embedding.visual.")))}
    for name, frames in cases.items():
        out["in_" + name] = frames
        out["emb_" + name] = np.atleast_2d(net(torch.from_numpy(frames)))
        print(name, out["emb_" + name].shape, float(np.abs(out["emb_" + name]).mean()))
    np.savez_compressed(os.path.join(GOLDEN, "clip_rn50.npz"), **out)


def golden_clip_transforms(E):
    """The `transforms` the reference builds for 'clip_vit' (src/embeddings.py:309-314: antialiased bicubic Resize(224)
    -> CenterCrop(224) -> float -> CLIP Normalize), run on frames that are not 224x224 (Habitat renders 64x64). The
    `clip` package is not installed: `clip.load` is replaced by a stand-in that only carries `visual.input_resolution`,
    which is all `_get_embedding` reads before building the transforms; the network itself is not used here."""
    import types

    class FakeClipModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = types.SimpleNamespace(input_resolution=224)
            self.dummy = torch.nn.Parameter(torch.zeros(1))

    fake = types.ModuleType("clip")
    fake.load = lambda name, device="cpu": (FakeClipModel(), None)
    old = getattr(E, "clip", None)
    E.clip = fake
    try:
        _, tf = E._get_embedding("clip_vit")
    finally:
        if old is not None:
            E.clip = old
    resize_crop = torch.nn.Sequential(tf[0], tf[1])
    to_float = torch.nn.Sequential(tf[2], tf[3])
    cases = {
        "structured_64": restate.structured_frames(2, 64, 64, 3, 51),
        "structured_96x128": restate.structured_frames(1, 96, 128, 3, 52),
        "noise_100x75": np.random.default_rng(53).integers(0, 256, (1, 100, 75, 3), dtype=np.uint8),
        "structured_336x448": restate.structured_frames(1, 336, 448, 3, 54),
        "adversarial_64": restate.adversarial_frames(64, 64),
        "noise_224": np.random.default_rng(55).integers(0, 256, (1, 224, 224, 3), dtype=np.uint8),
    }
    out = {}
    for name, frames in cases.items():
        x = torch.from_numpy(frames).permute(0, 3, 1, 2).contiguous()
        u = resize_crop(x)
        assert u.dtype == torch.uint8 and tuple(u.shape[2:]) == (224, 224)
        out["in_" + name] = frames
        out["u8_" + name] = u.numpy()
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).repeat(1, 3, 1, 1)
    out["lut"] = to_float(ramp).numpy().reshape(3, 256)
    np.savez_compressed(os.path.join(GOLDEN, "clip_transforms.npz"), **out)
    print("clip_transforms.npz:", {k: v.shape for k, v in out.items() if k.startswith("u8_")})


def golden_save_embedded(E):
    """behavioral_cloning/save_embedded_obs.py `run(flags)` UNMODIFIED, both sources, on three synthetic ImageNav-style
    trajectories (64x64, current || goal = 6 channels) with the 'random' encoder on CPU: the file names, pickle layouts
    and values the drop-in pvr_habitat_b200/save_embedded_obs.py has to reproduce."""
    import importlib
    import pickle
    import cv2
    S = importlib.import_module("behavioral_cloning.save_embedded_obs")
    lengths = [5, 3, 4]
    rng = np.random.default_rng(31)
    traj = dict(obs=[], action=[], reward=[], done=[], true_state=[])
    for i, n in enumerate(lengths):
        frames = restate.structured_frames(n, 64, 64, 6, 310 + i)
        frames[:, :, :, 3:] = frames[-1:, :, :, 3:]          # one goal image per trajectory, as ImageNav writes it
        traj["obs"].append(frames)
        traj["action"].append(rng.integers(0, 3, n))
        traj["reward"].append(rng.random(n).astype(np.float32))
        d = np.zeros(n, dtype=bool)
        d[-1] = True
        traj["done"].append(d)
        traj["true_state"].append(rng.standard_normal((n, 12)).astype(np.float32))
    out = {"lengths": np.array(lengths)}
    for k, v in traj.items():
        out["in_" + k] = np.concatenate(v)
    env, run_id = "HabitatImageNav-apartment_0", 4
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, env + ".pickle"), "wb") as fh:
            pickle.dump(traj, fh, protocol=pickle.HIGHEST_PROTOCOL)
        os.makedirs(os.path.join(d, env))
        for t, frames in enumerate(traj["obs"]):             # save_opt_trajectories_png.py:43-58
            for s_ in range(len(frames)):
                cv2.imwrite(os.path.join(d, env, f"{t}_{s_}.png"), frames[s_][:, :, :3])
            cv2.imwrite(os.path.join(d, env, f"{t}_goal.png"), frames[-1][:, :, 3:])
            with open(os.path.join(d, env, f"{t}.pickle"), "wb") as fh:
                pickle.dump({k: traj[k][t] for k in ("action", "reward", "done", "true_state")}, fh,
                            protocol=pickle.HIGHEST_PROTOCOL)
        for source in ("pickle", "png"):
            flags = S.parser.parse_args(["--data_path", d, "--env", env, "--embedding_name", "random", "--run_id",
                                         str(run_id), "--batch_size", "4", "--disable_cuda", "--source", source,
                                         "--n_trajectories", "-1"])
            S.run(flags)
            name = os.path.join(d, env + "_random.pickle")
            with open(name, "rb") as fh:
                data = pickle.load(fh)
            os.remove(name)
            out[f"{source}_keys"] = np.array(list(data.keys()))
            for k, v in data.items():
                out[f"{source}_{k}"] = np.array([os.path.relpath(x, d) for x in v]) if k == "png" else np.asarray(v)
            ck = torch.load(os.path.join(d, f"random_{run_id}.tar"), map_location="cpu")
            out["tar_keys"] = np.array(list(ck.keys()))
            for k, v in ck["embedding_model_state_dict"].items():
                out["w_" + k] = v.numpy()
            out["files"] = np.array(sorted(f for f in os.listdir(d) if os.path.isfile(os.path.join(d, f))))
    out["env"], out["run_id"] = np.array(env), np.array(run_id)
    np.savez_compressed(os.path.join(GOLDEN, "save_embedded.npz"), **out)
    print("save_embedded.npz:", {k: v.shape for k, v in out.items() if k.startswith(("pickle_", "png_"))})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["transforms", "embeddings", "policy", "small_conv"]
    if "transforms" in which or "embeddings" in which or "small_conv" in which or "resnet_basic" in which or \
            "mae" in which or "mae_huge" in which or "maskrcnn" in which or "clip_rn50" in which or "save_embedded" in which or "clip_transforms" in which:
        E = refshim.reference_embeddings()
        if "clip_transforms" in which:
            golden_clip_transforms(E)
        if "save_embedded" in which:
            golden_save_embedded(E)
        if "mae" in which:
            golden_mae(E)
        if "mae_huge" in which:
            golden_mae_huge(E)
        if "maskrcnn" in which:
            golden_maskrcnn(E)
        if "clip_rn50" in which:
            golden_clip_rn50(E)
        if "resnet_basic" in which:
            golden_resnet_basic(E)
        if "transforms" in which:
            golden_transforms(E)
        if "embeddings" in which:
            golden_embeddings(E)
        if "small_conv" in which:
            golden_small_conv(E)
    if "policy" in which:
        golden_policy()


if __name__ == "__main__":
    main()
