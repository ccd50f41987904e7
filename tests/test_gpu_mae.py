"""-m gpu: the MAE path (bicubic preprocessing kernel, erf-GELU / width-1024 ViT kernels, whole encoders) through the
C ABI against the oracle (oracle/restate.py, oracle/restate_mae.py) and the reference goldens (tests/golden/mae.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, restate_mae
from pvr_habitat_b200 import _lib, models
from pvr_habitat_b200.embeddings import EmbeddingNet, Transforms
from pvr_habitat_b200.vision_models.moco import allow_random_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    """mae.npz (mae_base / mae_large) + mae_huge.npz (the first frame of each frame set), one key space."""
    g = dict(np.load(os.path.join(golden_dir, "mae.npz")))
    huge = np.load(os.path.join(golden_dir, "mae_huge.npz"))
    g.update({k: huge[k] for k in huge.files if k.endswith("mae_huge")})
    return g


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cuda_bicubic(frames_nhwc, nf=1):
    t = Transforms(interpolation="bicubic")
    n = frames_nhwc.shape[0]
    out = torch.full((nf * n, 3, 224, 224), float("nan"), device="cuda")
    t.run(torch.from_numpy(frames_nhwc).cuda(), nf, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ------------------------------------------------------------------------------------------------ K1, bicubic: bit exact
@pytest.mark.parametrize("case", ["structured_64", "structured_224", "structured_96x128", "noise_224", "adversarial_64"])
def test_bicubic_preprocess_bit_exact_vs_reference_golden(gold, case):
    lut = restate.normalize_lut()
    u = gold["u8_" + case]
    want = np.stack([lut[c][u[:, c]] for c in range(3)], 1)
    got = cuda_bicubic(gold["in_" + case])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{int((got != want).sum())} values differ"


@pytest.mark.parametrize("hw,nf,n", [((100, 75), 1, 3), ((33, 47), 2, 2), ((480, 640), 1, 1), ((84, 84), 3, 2)])
def test_bicubic_preprocess_bit_exact_vs_oracle_ragged_shapes(hw, nf, n):
    """Non-dyadic ratios, down-scaling, band borders at the image edge, several frames per observation."""
    obs = np.random.default_rng(hw[0] * 7 + nf).integers(0, 256, (n, hw[0], hw[1], 3 * nf), dtype=np.uint8)
    frames, _ = restate.split_frames(obs)
    want = restate.transforms(np.ascontiguousarray(np.transpose(frames, (0, 3, 1, 2))), interpolation="bicubic")
    got = cuda_bicubic(obs, nf)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{int((got != want).sum())} values differ"


def test_bicubic_bf16_layout_is_the_rounded_float_output():
    obs = restate.structured_frames(3, 64, 64, 6, 5)
    t = Transforms(interpolation="bicubic")
    out = torch.zeros(6, 224, 224, 4, dtype=torch.bfloat16, device="cuda")
    t.run(torch.from_numpy(obs).cuda(), 2, out.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    f32 = cuda_bicubic(obs, 2)  # frame-major
    want = torch.from_numpy(f32).reshape(2, 3, 3, 224, 224).permute(1, 0, 3, 4, 2).reshape(6, 224, 224, 3)
    assert torch.equal(out[..., :3].cpu(), want.to(torch.bfloat16)) and float(out[..., 3].abs().max()) == 0


# ------------------------------------------------------------------------------------------------ kernels
def test_gemm_erf_gelu_epilogue():
    g = torch.Generator().manual_seed(3)
    m, k, n = 777, 1024, 4096
    a = torch.randn(m, k, generator=g).bfloat16().cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().cuda()
    b = torch.randn(n, generator=g).cuda()
    h = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    models.gemm(a, w, h, m, n, k, bias=b, act=3)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + b)
    assert rel(h.float(), ref) < 4e-3  # bf16 output rounding


@pytest.mark.parametrize("width", [768, 1024, 1280])
def test_layernorm_f32_and_embed_without_ln(width):
    g = torch.Generator().manual_seed(width)
    lib, st = _lib.lib(), _lib.current_stream_ptr
    x = (torch.randn(40 * 5, width, generator=g) * 2 + 0.5).cuda()
    w, b = torch.randn(width, generator=g).cuda(), torch.randn(width, generator=g).cuda()
    y = torch.empty(40, width, device="cuda")
    _lib.check(lib.pvr_layernorm_f32(x.data_ptr(), 5, 40, width, w.data_ptr(), b.data_ptr(), 1e-6, y.data_ptr(), width,
                                     st()))
    ref = torch.nn.functional.layer_norm(x[::5], (width,), w, b, 1e-6)
    assert rel(y, ref) < 1e-6
    ybf = torch.empty(40 * 5, width, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.pvr_layernorm(x.data_ptr(), 1, 200, width, w.data_ptr(), b.data_ptr(), 1e-6, ybf.data_ptr(), st()))
    assert rel(ybf.float(), torch.nn.functional.layer_norm(x, (width,), w, b, 1e-6)) < 3e-3
    # token assembly without ln_pre: [cls | patches] + pos
    n, tokens = 3, 17
    patches = torch.randn(n * (tokens - 1), width, generator=g).bfloat16().cuda()
    cls, pos = torch.randn(width, generator=g).cuda(), torch.randn(tokens, width, generator=g).cuda()
    out = torch.empty(n * tokens, width, device="cuda")
    _lib.check(lib.pvr_vit_embed(patches.data_ptr(), cls.data_ptr(), pos.data_ptr(), n, tokens, width, None, None, 0.0,
                                 out.data_ptr(), st()))
    want = torch.cat([cls.expand(n, 1, width), patches.float().reshape(n, tokens - 1, width)], 1) + pos
    assert torch.equal(out.reshape(n, tokens, width), want)


def test_attention_16_heads():
    tokens, n_img, W, H = 197, 2, 1024, 16
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(n_img * tokens, 3 * W, generator=g).bfloat16().cuda()
    out = torch.full((n_img * tokens, W), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().pvr_attention(qkv.data_ptr(), n_img, tokens, W, H, out.data_ptr(), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = (t.float().reshape(n_img, tokens, H, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(n_img * tokens, W)
    assert not torch.isnan(out.float()).any() and rel(out.float(), ref) < 1e-2


@pytest.mark.parametrize("tokens,n_img,H,D", [(257, 3, 16, 80), (197, 2, 12, 64), (50, 5, 4, 128), (33, 2, 3, 96),
                                               (300, 1, 2, 80)])
def test_attention_mma_any_head_dim(tokens, n_img, H, D):
    """csrc/attention_mma.cu (mae_huge: 257 tokens, 16 heads of 80) against float64 softmax attention on the same bf16
    inputs; at head_dim 64 it also has to agree with the tcgen05 kernel it stands in for."""
    W = H * D
    g = torch.Generator().manual_seed(tokens + D)
    qkv = torch.randn(n_img * tokens, 3 * W, generator=g).bfloat16().cuda()
    out = torch.full((n_img * tokens, W), float("nan"), dtype=torch.bfloat16, device="cuda")
    lib = _lib.lib()
    _lib.check(lib.pvr_attention_mma(qkv.data_ptr(), n_img, tokens, W, H, out.data_ptr(), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = (t.double().reshape(n_img, tokens, H, D).transpose(1, 2) for t in qkv.chunk(3, -1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * D ** -0.5, -1) @ v).transpose(1, 2).reshape(n_img * tokens, W)
    assert not torch.isnan(out.float()).any() and rel(out.float(), ref) < 6e-3
    via = torch.full_like(out, float("nan"))  # the public entry dispatches on the shape
    _lib.check(lib.pvr_attention(qkv.data_ptr(), n_img, tokens, W, H, via.data_ptr(), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    if D == 64 and tokens <= 256:
        assert rel(via.float(), out.float()) < 6e-3  # tensor-memory kernel
    else:
        assert torch.equal(via, out)


# ------------------------------------------------------------------------------------------------ encoders
# north star: bf16 embeddings within relative L2 <= 1e-2 and cosine >= 0.999 of the reference
def check_embedding(got, ref):
    got, ref = np.atleast_2d(got).astype(np.float64), np.atleast_2d(ref).astype(np.float64)
    assert got.shape == ref.shape
    r = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert r <= 1e-2 and cos.min() >= 0.999, (r, cos.min())
    return r


def make_net(name, seed):
    with allow_random_init():
        net = EmbeddingNet(name)
    net.embedding.load_state_dict(restate_mae.mae_state(name, seed), strict=False)
    net.invalidate()
    return net


@pytest.mark.parametrize("name", ["mae_base", "mae_large", "mae_huge"])
def test_mae_embedding_vs_reference_golden(gold, name):
    net = make_net(name, int(gold[f"seed_{name}"]))
    assert net.out_size == int(gold[f"out_size_{name}"])
    for tag in ("64", "224"):
        ref = gold[f"emb{tag}_{name}"]  # mae_huge: the first frame only
        got = net(torch.from_numpy(gold["frames" + tag][:len(ref)]))
        assert isinstance(got, np.ndarray) and got.dtype == np.float32
        r = check_embedding(got, ref)
        print(f"{name} {tag}: rel-L2 {r:.2e}")


@pytest.mark.parametrize("name", ["mae_base", "mae_large", "mae_huge"])
def test_mae_embedding_fp32_mode_vs_reference_golden(gold, name):
    """North star: embeddings within relative L2 <= 1e-5 "in the fp32 mode" (net.set_precision('fp32'): float32
    weights, activations and accumulation on the CUDA cores, csrc/vit_f32.cu)."""
    net = make_net(name, int(gold[f"seed_{name}"])).set_precision('fp32')
    for tag in ("64", "224"):
        ref = np.atleast_2d(gold[f"emb{tag}_{name}"]).astype(np.float64)
        got = np.atleast_2d(net(torch.from_numpy(gold["frames" + tag][:len(ref)]))).astype(np.float64)
        r = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"{name} {tag} fp32 mode: rel-L2 {r:.2e}")
        assert r <= 1e-5, r


def test_mae_two_frame_observation_and_batch_independence(gold):
    net = make_net("mae_base", int(gold["seed_mae_base"]))
    obs = restate.structured_frames(5, 64, 64, 6, 61)
    fused = net.embed(torch.from_numpy(obs), n_frames=2).cpu().numpy()
    assert fused.shape == (5, 2 * 768)
    frames, _ = restate.split_frames(obs)
    host = restate.regroup_frames(net(torch.from_numpy(frames)), 2)
    assert np.array_equal(host, fused)
    sd = restate_mae.mae_state("mae_base", int(gold["seed_mae_base"]))
    check_embedding(host[:2], restate.regroup_frames(restate_mae.embedding_forward(sd, "mae_base", frames[[0, 1, 5, 6]]), 2))
    single = net(torch.from_numpy(frames[:1]))
    assert single.shape == (768,)  # squeezed like the reference (src/embeddings.py:402)
    check_embedding(single, host[0, :768])
