"""-m gpu: the CUDA path (through the C ABI) against the oracle and the reference-generated golden fixtures."""
import os

import numpy as np
import pytest
import torch

import gpu_cases
from oracle import restate
from pvr_habitat_b200 import _lib
from pvr_habitat_b200.embeddings import EmbeddingNet, Transforms
from pvr_habitat_b200.vision_models.moco import allow_random_init

pytestmark = pytest.mark.gpu

VARIANTS = {"moco_aug": ["conv5"], "moco_aug_l4": ["l4"], "moco_aug_l3": ["l3"], "moco_aug_uber_34": ["l3", "l4"],
            "moco_aug_uber_345": ["l3", "l4", "conv5"]}


@pytest.fixture(scope="module")
def tf(golden_dir):
    return np.load(os.path.join(golden_dir, "transforms.npz"))


@pytest.fixture(scope="module")
def emb(golden_dir):
    return np.load(os.path.join(golden_dir, "embeddings.npz"))


def make_net(name, seeds):
    seed = {"conv5": int(seeds[0]), "l4": int(seeds[1]), "l3": int(seeds[2])}
    with allow_random_init():
        net = EmbeddingNet(name)
    parts = net.embedding.models if hasattr(net.embedding, "models") else [net.embedding]
    for m, v in zip(parts, VARIANTS[name]):
        m.load_state_dict(restate.resnet50_state(v, seed[v]), strict=True)
    net.invalidate()
    return net


def cuda_transforms(frames_nhwc, nf=1):
    t = Transforms()
    n = frames_nhwc.shape[0]
    out = torch.full((nf * n, 3, 224, 224), float("nan"), device="cuda")
    t.run(torch.from_numpy(frames_nhwc).cuda(), nf, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ------------------------------------------------------------------------------------------------ K1 (bit exact)
@pytest.mark.parametrize("case", ["structured_64", "structured_224", "noise_64", "noise_224", "adversarial_64",
                                  "adversarial_224"])
def test_preprocess_bit_exact_vs_reference_golden(tf, case):
    got = cuda_transforms(tf["in_" + case])
    u, lut = tf["u8_" + case], tf["lut"]
    ref = np.stack([lut[c][u[:, c]] for c in range(3)], 1)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))  # 0 ulp (north star allows 1)


@pytest.mark.parametrize("hw,nf,n", [((64, 64), 2, 7), ((96, 128), 3, 3), ((100, 75), 1, 2), ((480, 640), 1, 2),
                                     ((84, 84), 4, 1), ((224, 224), 3, 2), ((33, 47), 1, 3)])
def test_preprocess_bit_exact_vs_oracle_ragged_shapes(hw, nf, n):
    rng = np.random.default_rng(hw[0] * 7 + nf)
    obs = rng.integers(0, 256, (n, hw[0], hw[1], 3 * nf), dtype=np.uint8)
    got = cuda_transforms(obs, nf)
    frames, _ = restate.split_frames(obs)  # frame-major like the reference's host loop
    ref = restate.transforms(np.ascontiguousarray(frames.transpose(0, 3, 1, 2)))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_preprocess_bf16_nhwc4_sample_major():
    obs = restate.structured_frames(3, 64, 64, 6, 5)
    t = Transforms()
    out = torch.full((6, 224, 224, 4), float("nan"), dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, out.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    frames, _ = restate.split_frames(obs)
    ref = torch.from_numpy(restate.transforms(np.ascontiguousarray(frames.transpose(0, 3, 1, 2))))
    ref = ref.view(2, 3, 3, 224, 224).permute(1, 0, 3, 4, 2).reshape(6, 224, 224, 3).to(torch.bfloat16)
    got = out.cpu()
    assert torch.equal(got[..., :3], ref) and torch.all(got[..., 3] == 0)


def test_preprocess_stem_layout_sample_major():
    """PVR_FMT_STEM_BF16: per stem output column q, the 8 input columns 2q-3..2q+4 x 4 channels (zeros outside)."""
    from pvr_habitat_b200.program import expand_stem_input
    obs = restate.structured_frames(3, 64, 64, 6, 6)
    t = Transforms()
    out = torch.full((6, 224, 112, 32), float("nan"), dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, out.data_ptr(), _lib.PVR_FMT_STEM_BF16, True)
    frames, _ = restate.split_frames(obs)
    ref = torch.from_numpy(restate.transforms(np.ascontiguousarray(frames.transpose(0, 3, 1, 2))))
    ref = ref.view(2, 3, 3, 224, 224).permute(1, 0, 3, 4, 2).reshape(6, 224, 224, 3).to(torch.bfloat16)
    ref4 = torch.zeros(6, 224, 224, 4, dtype=torch.bfloat16)
    ref4[..., :3] = ref
    assert torch.equal(out.cpu(), expand_stem_input(ref4))


def test_preprocess_compact_stem_layout_and_bitwise_equal_embeddings(emb, monkeypatch):
    """PVR_FMT_STEM_PAD_BF16: padded NHWC4 rows (3 zero pixels before column 0, 5 after the last), which the stem's
    tensor map expands into the 8-column windows of PVR_FMT_STEM_BF16 (overlapping 64-byte boxes 16 bytes apart).
    (1) the layout, (2) embeddings bitwise equal to the ones computed from the materialised W-expanded input."""
    obs = restate.structured_frames(3, 64, 64, 6, 6)
    t = Transforms()
    out = torch.full((6, 224, 232, 4), float("nan"), dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, out.data_ptr(), _lib.PVR_FMT_STEM_PAD_BF16, True)
    nhwc = torch.full((6, 224, 224, 4), float("nan"), dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, nhwc.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    assert torch.equal(out[:, :, 3:227], nhwc)
    assert float(out[:, :, :3].abs().max()) == 0 and float(out[:, :, 227:].abs().max()) == 0
    frames = restate.structured_frames(5, 224, 224, 3, 23)
    net = make_net("moco_aug_uber_34", emb["weight_seeds"])
    assert net.encoder().input_format == _lib.PVR_FMT_STEM_PAD_BF16
    compact = net.embed(torch.from_numpy(frames)).cpu().numpy()
    monkeypatch.setenv("PVR_STEM_EXPANDED", "1")
    net2 = make_net("moco_aug_uber_34", emb["weight_seeds"])
    assert net2.encoder().input_format == _lib.PVR_FMT_STEM_BF16
    assert np.array_equal(compact, net2.embed(torch.from_numpy(frames)).cpu().numpy())
    monkeypatch.setenv("PVR_NO_POOL_FUSION", "1")  # the un-fused stem (tile origin without the pooling halo)
    monkeypatch.delenv("PVR_STEM_EXPANDED")
    net3 = make_net("moco_aug_uber_34", emb["weight_seeds"])
    assert np.array_equal(compact, net3.embed(torch.from_numpy(frames)).cpu().numpy())


def test_transforms_module_keeps_reference_calling_convention(tf):
    frames = tf["in_structured_64"]
    x = torch.from_numpy(frames).permute(0, 3, 1, 2).contiguous().cuda()  # NCHW uint8, as src/embeddings.py:392-393
    got = Transforms()(x).cpu().numpy()
    ref = restate.transforms(np.ascontiguousarray(frames.transpose(0, 3, 1, 2)))
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------------------------------------ GEMM / conv
# bf16 inputs, fp32 accumulate, one bf16 rounding of the output: rel-L2 vs the fp32 computation ~ 2^-9 / sqrt(3)
GEMM_TOL = 3e-3


@pytest.mark.parametrize("m,n,k,relu,res", [(128, 64, 64, False, False), (1, 64, 64, True, False),
                                            (129, 128, 192, True, True), (1000, 256, 512, True, True),
                                            (5000, 1024, 1024, False, True), (300, 4096, 1024, False, False)])
def test_gemm_vs_fp32(m, n, k, relu, res):
    r = gpu_cases.case_gemm(m, n, k, relu, res)
    assert r["nan"] == 0 and r["rel_l2"] < GEMM_TOL, r


@pytest.mark.parametrize("args", [
    dict(n=3, h=14, wd=14, ci=128, co=256, r=1, stride=1, pad=0, relu=True, res=True),
    dict(n=3, h=14, wd=14, ci=128, co=128, r=3, stride=1, pad=1),
    dict(n=2, h=56, wd=56, ci=64, co=64, r=3, stride=1, pad=1),           # patch-resident kernel
    dict(n=3, h=24, wd=16, ci=64, co=64, r=3, stride=1, pad=1, relu=False),  # patch kernel, clipped row tiles
    dict(n=1, h=8, wd=8, ci=64, co=64, r=3, stride=1, pad=1),               # patch kernel, single partial tile
    dict(n=3, h=28, wd=28, ci=128, co=128, r=3, stride=2, pad=1),
    dict(n=3, h=28, wd=28, ci=256, co=512, r=1, stride=2, pad=0, relu=False),
    dict(n=5, h=7, wd=7, ci=512, co=512, r=3, stride=1, pad=1, res=True),
    dict(n=1, h=7, wd=7, ci=2048, co=512, r=1, stride=1, pad=0),
    dict(n=37, h=14, wd=14, ci=256, co=256, r=3, stride=1, pad=1, block_n=256),
    dict(n=37, h=14, wd=14, ci=256, co=256, r=3, stride=1, pad=1, block_n=64),
])
def test_conv_vs_fp32(args):
    r = gpu_cases.case_conv(**args)
    assert r["nan"] == 0 and r["rel_l2"] < GEMM_TOL, r


def test_stem_conv_vs_fp32():
    r = gpu_cases.case_stem(2, 224)
    assert r["nan"] == 0 and r["rel_l2"] < GEMM_TOL, r
    r = gpu_cases.case_stem(3, 64)
    assert r["nan"] == 0 and r["rel_l2"] < GEMM_TOL, r


# ------------------------------------------------------------------------------------------------ embeddings
# north star: bf16 embeddings within relative L2 <= 1e-2 and cosine >= 0.999 of the reference
def check_embedding(got, ref):
    got, ref = np.atleast_2d(got).astype(np.float64), np.atleast_2d(ref).astype(np.float64)
    assert got.shape == ref.shape
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert rel <= 1e-2 and cos.min() >= 0.999, (rel, cos.min())
    return rel


@pytest.mark.parametrize("name", list(VARIANTS))
def test_embedding_vs_reference_golden(emb, name):
    net = make_net(name, emb["weight_seeds"])
    assert net.out_size == int(emb[f"out_size_{name}"])
    got = net(torch.from_numpy(emb["frames64"]))
    assert isinstance(got, np.ndarray) and got.dtype == np.float32
    check_embedding(got, emb[f"emb64_{name}"])
    if f"emb224_{name}" in emb:
        check_embedding(net(torch.from_numpy(emb["frames224"])), emb[f"emb224_{name}"])


@pytest.mark.parametrize("name", ["moco_aug", "moco_aug_l3"])
def test_two_frame_observation_fused_split_regroup(emb, name):
    """(N,H,W,6) observation: frame split + regroup of main_bc_1.py:128-136 done inside the kernels."""
    net = make_net(name, emb["weight_seeds"])
    obs = emb["obs2"]
    fused = net.embed(torch.from_numpy(obs), n_frames=2).cpu().numpy()
    check_embedding(fused, emb[f"emb_obs2_{name}"])
    # and the reference's own host-side loop on top of EmbeddingNet.__call__ gives the same numbers
    o = np.concatenate(np.split(obs, 2, axis=3), axis=0)
    o = net(torch.from_numpy(o))
    host = np.concatenate(np.split(o, 2, axis=0), axis=-1)
    assert np.array_equal(host, fused)


def test_single_frame_is_squeezed_like_reference(emb):
    net = make_net("moco_aug", emb["weight_seeds"])
    out = net(torch.from_numpy(emb["frames64"][:1]))
    assert out.shape == (2048,)  # src/embeddings.py:402 .squeeze()
    check_embedding(out, emb["emb64_moco_aug"][:1])


def test_full_size_properties(emb):
    """BASELINE-size batch (256 frames of 224x224): size-independent properties instead of an oracle run —
    permutation equivariance (bitwise, the kernels are deterministic), batch-composition independence, and a
    spot check of 4 frames against the oracle."""
    net = make_net("moco_aug", emb["weight_seeds"])
    frames = restate.structured_frames(32, 224, 224, 3, 99)
    big = np.concatenate([frames] * 8)  # 256 frames
    perm = np.random.default_rng(0).permutation(256)
    a = net(torch.from_numpy(big))
    b = net(torch.from_numpy(big[perm]))
    assert np.array_equal(a[perm], b)
    assert np.array_equal(a[:32], a[32:64])  # same frame, different tile position -> same bits
    small = net(torch.from_numpy(frames[:5]))  # ragged M tail (5*49 rows in layer4)
    assert np.array_equal(small, a[:5])
    parts = [("conv5", restate.resnet50_state("conv5", int(emb["weight_seeds"][0])))]
    check_embedding(a[:4], restate.embedding_forward(parts, frames[:4]))


def test_load_state_dict_recompiles(emb):
    net = make_net("moco_aug", emb["weight_seeds"])
    x = torch.from_numpy(emb["frames64"])
    a = net(x)
    sd = net.state_dict()
    sd["embedding.layer4.2.bn3.weight"] = sd["embedding.layer4.2.bn3.weight"] * 2
    net.load_state_dict(sd)
    b = net(x)
    assert not np.array_equal(a, b)


def test_random_small_conv_embedding_vs_reference_golden(golden_dir):
    """'random' PVR (5 x conv3x3/s2 + ELU) against the reference's EmbeddingNet('random') outputs."""
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    torch.manual_seed(9)
    net = EmbeddingNet("random", pretrained=False)
    net.embedding.load_state_dict({k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w_")})
    net.invalidate()
    assert net.out_size == 1568
    for key in ("64", "224"):
        got = net(torch.from_numpy(gold["frames" + key]))
        check_embedding(got, gold["emb" + key])


def test_long_observation_arrays_are_embedded_in_passes(emb):
    """More images than one pass holds (workspace bound): rows identical to the single-pass result."""
    net = make_net("moco_aug", emb["weight_seeds"])
    frames = restate.structured_frames(11, 64, 64, 3, 77)
    whole = net.embed(torch.from_numpy(frames)).cpu().numpy()
    net.max_images_per_pass = 4  # 3 passes: 4 + 4 + 3 (ragged)
    parts = net.embed(torch.from_numpy(frames)).cpu().numpy()
    assert np.array_equal(whole, parts)


def test_fused_stem_maxpool_and_zigzag_are_bit_identical_to_the_plain_schedule(emb, monkeypatch):
    """The max pool fused into the stem epilogue (conv3x3_patch.cu MODE 2), the zig-zag tile order, programmatic
    dependent launch and the CTA-pair (cta_group::2) kernels only change the schedule: embeddings are bitwise those of
    the separate stem -> maxpool3x3s2_kernel, first-to-last order, one CTA per tile, and of conv3 / next conv1 as two
    kernels instead of the back-to-back fused one (conv_b2b.cu). 128 frames so that the pair
    kernels are selected (they need >= 74 pair tiles)."""
    frames = restate.structured_frames(128, 64, 64, 3, 5)
    net = make_net("moco_aug", emb["weight_seeds"])
    fused = net.embed(torch.from_numpy(frames)).cpu().numpy()
    monkeypatch.setenv("PVR_NO_POOL_FUSION", "1")
    monkeypatch.setenv("PVR_NO_ZIGZAG", "1")
    monkeypatch.setenv("PVR_NO_PDL", "1")
    monkeypatch.setenv("PVR_CTA2", "0")
    monkeypatch.setenv("PVR_NO_B2B", "1")
    net2 = make_net("moco_aug", emb["weight_seeds"])
    plain = net2.embed(torch.from_numpy(frames)).cpu().numpy()
    assert np.array_equal(fused, plain)


def test_fused_kernels_bit_identical_on_ragged_tiles(emb, monkeypatch):
    """3 frames of 224 x 224: 9408 layer1 pixels = 73.5 tiles of 128 rows, so the back-to-back kernel, the fused
    stem / max pool and the projection-shortcut GEMM all see a ragged last tile (TMA zero fill / clipping). Same
    bitwise comparison against the plain schedule as above, on the two-trunk model with compression heads."""
    frames = restate.structured_frames(3, 224, 224, 3, 11)
    net = make_net("moco_aug_uber_34", emb["weight_seeds"])
    fused = net.embed(torch.from_numpy(frames)).cpu().numpy()
    for k, v in (("PVR_NO_POOL_FUSION", "1"), ("PVR_NO_ZIGZAG", "1"), ("PVR_NO_PDL", "1"), ("PVR_CTA2", "0"),
                 ("PVR_NO_B2B", "1")):
        monkeypatch.setenv(k, v)
    net2 = make_net("moco_aug_uber_34", emb["weight_seeds"])
    plain = net2.embed(torch.from_numpy(frames)).cpu().numpy()
    assert np.array_equal(fused, plain)


def test_streamed_back_to_back_kernel_is_bit_identical(emb, monkeypatch):
    """conv_b2b_stream_kernel (layer2: 128 -> 512 + residual fused with the next 512 -> 128, weights streamed through
    rings; the default since the end of round 2, PVR_B2B_STREAM=0 turns it off): its embeddings are bitwise those of
    the two separate kernels."""
    frames = restate.structured_frames(6, 224, 224, 3, 17)
    monkeypatch.setenv("PVR_B2B_STREAM", "0")
    net = make_net("moco_aug", emb["weight_seeds"])
    separate = net.embed(torch.from_numpy(frames)).cpu().numpy()
    monkeypatch.delenv("PVR_B2B_STREAM")
    net2 = make_net("moco_aug", emb["weight_seeds"])
    streamed = net2.embed(torch.from_numpy(frames)).cpu().numpy()
    assert np.array_equal(separate, streamed)
    assert net2.encoder().lib.pvr_encoder_launch_count(net2.encoder().handle) < \
        net.encoder().lib.pvr_encoder_launch_count(net.encoder().handle)


@pytest.mark.parametrize("name", ["resnet18", "resnet34"])
def test_resnet_basic_embedding_vs_reference_golden(golden_dir, name):
    """`resnet18` / `resnet34` (BasicBlock nets, src/embeddings.py:112-117) against the reference's
    EmbeddingNet outputs on the same synthetic weights."""
    g = np.load(os.path.join(golden_dir, "resnet_basic.npz"))
    with allow_random_init():
        net = EmbeddingNet(name, pretrained=False)
    net.embedding.load_state_dict(restate.resnet_basic_state(name, int(g[f"seed_{name}"])), strict=True)
    net.invalidate()
    assert net.out_size == 512
    for tag in ("64", "224"):
        check_embedding(net(torch.from_numpy(g["frames" + tag])), g[f"emb{tag}_{name}"])


def test_small_batch_graph_replay_equals_eager(emb):
    """Rollout-sized batches (<= 8 images) into a fixed output buffer: the first call runs eagerly, the second is
    captured into a CUDA graph, later ones replay it — all bitwise equal, also after the input changes."""
    net = make_net("moco_aug_uber_34", emb["weight_seeds"])
    out = torch.empty(2, net.out_size, device="cuda")
    ref_net = make_net("moco_aug_uber_34", emb["weight_seeds"])
    for k in range(5):
        frames = torch.from_numpy(restate.structured_frames(2, 64, 64, 3, 30 + k))
        net.embed(frames, 1, out)
        want = ref_net.embed(frames)  # fresh output tensor every time: never replayed
        assert torch.equal(out, want), k


# ------------------------------------------------------------------------------------------------ fp32 parity mode
# north star: embeddings within relative L2 <= 1e-5 of the reference "in the fp32 mode" (EmbeddingNet.set_precision)
FP32_TOL = 1e-5


def check_embedding_fp32(got, ref):
    got, ref = np.atleast_2d(got).astype(np.float64), np.atleast_2d(ref).astype(np.float64)
    assert got.shape == ref.shape
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    rows = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel <= FP32_TOL and rows.max() <= FP32_TOL, (rel, rows.max())
    return rel


@pytest.mark.parametrize("name", list(VARIANTS))
def test_fp32_mode_embedding_vs_reference_golden(emb, name):
    net = make_net(name, emb["weight_seeds"]).set_precision("fp32")
    check_embedding_fp32(net(torch.from_numpy(emb["frames64"])), emb[f"emb64_{name}"])
    if f"emb224_{name}" in emb:
        check_embedding_fp32(net(torch.from_numpy(emb["frames224"])), emb[f"emb224_{name}"])


def test_fp32_mode_two_frame_observation_and_switching_back(emb):
    net = make_net("moco_aug_l3", emb["weight_seeds"])
    obs = torch.from_numpy(emb["obs2"])
    bf16 = net.embed(obs, n_frames=2).cpu().numpy()
    fp32 = net.set_precision("fp32").embed(obs, n_frames=2).cpu().numpy()
    check_embedding_fp32(fp32, emb["emb_obs2_moco_aug_l3"])
    check_embedding(bf16, fp32)
    again = net.set_precision("bf16").embed(obs, n_frames=2).cpu().numpy()
    assert np.array_equal(again, bf16)  # the tensor-core program is rebuilt, bit for bit


@pytest.mark.parametrize("name", ["resnet18", "resnet34"])
def test_fp32_mode_resnet_basic_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "resnet_basic.npz"))
    with allow_random_init():
        net = EmbeddingNet(name, pretrained=False)
    net.embedding.load_state_dict(restate.resnet_basic_state(name, int(g[f"seed_{name}"])), strict=True)
    net.invalidate()
    net.set_precision("fp32")
    for tag in ("64", "224"):
        check_embedding_fp32(net(torch.from_numpy(g["frames" + tag])), g[f"emb{tag}_{name}"])


def test_fp32_mode_small_conv_vs_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "small_conv.npz"))
    net = EmbeddingNet("random", pretrained=False)
    net.embedding.load_state_dict({k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w_")})
    net.invalidate()
    net.set_precision("fp32")
    for key in ("64", "224"):
        check_embedding_fp32(net(torch.from_numpy(gold["frames" + key])), gold["emb" + key])


def test_a_stationary_tile_order_is_bitwise_neutral(monkeypatch):
    """ConvGemmParams::astat (layer3's 256 -> 1024 + residual convs walk all N tiles of an M tile back to back and reload
    only the weight chunks): another tile ORDER of the same arithmetic, so the embeddings are bit-identical. PVR_ASTAT=2
    forces it at a batch where the M-tile padding would normally rule it out (it is on by default from ~1200 frames)."""
    from pvr_habitat_b200.embeddings import EmbeddingNet
    from pvr_habitat_b200.vision_models.moco import allow_random_init
    frames = torch.from_numpy(restate.structured_frames(16, 64, 64, 3, 17))  # >= 13 frames: 128-wide tiles at layer3
    outs = []
    for mode in ("0", "2"):
        monkeypatch.setenv("PVR_ASTAT", mode)
        torch.manual_seed(4)
        with allow_random_init():
            net = EmbeddingNet("moco_aug_uber_34")
        outs.append(net(frames))
    assert outs[0].shape == (16, 4214) and np.isfinite(outs[0]).all()
    assert np.array_equal(outs[0], outs[1])
