"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm).

CPU restatement of the reference's PVR-embed -> BC-train hot path. Nothing under pvr_habitat_b200/ imports this.

The reference (sparisi/pvr_habitat) is pure Python on top of torchvision 0.10 / torch 1.9; its arithmetic for this
path is executed by those libraries, so each function below restates the *library* algorithm the reference calls and
cites both the reference call site and the library lines (tv: = torchvision 0.26 in this image, same algorithm).

Pinning: the reference ships no tests or golden vectors. The restatement is pinned against the
reference's own code executed in the build container (oracle/make_golden.py imports /root/reference unmodified and
writes tests/golden/*.npz); tests/test_oracle_golden.py re-checks it on every run.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


# ------------------------------------------------------------------------------------------------ K1: transforms
def resize_geometry(h, w, size=256, crop=224):
    """tv:transforms/functional.py:368-384 (_compute_resized_output_size, int size -> short side) and :592-594
    (center_crop offsets int(round((dim - crop) / 2.0)))."""
    if h <= w:
        rh, rw = size, int(size * w / h)
    else:
        rh, rw = int(size * h / w), size
    return rh, rw, int(round((rh - crop) / 2.0)), int(round((rw - crop) / 2.0))


def _fma(a, b, c):
    """float32 fused multiply-add emulated in float64 (the product of two float32 is exact in float64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _src_index(scale, dst, size):
    """ATen area_pixel_compute_source_index(align_corners=False) + guard_index_and_lambda in float32.
    The x86 (AVX2/FMA) build of ATen contracts scale*(dst+0.5)-0.5 into one FMA; probed bit-exact."""
    f32 = np.float32
    d = dst.astype(f32) + f32(0.5)
    s = _fma(np.full_like(d, f32(scale)), d, np.full_like(d, f32(-0.5)))
    s = np.maximum(s, f32(0))
    i0 = np.minimum(s.astype(np.int64), size - 1)
    i1 = np.minimum(i0 + 1, size - 1)
    lam = np.clip(s - i0.astype(f32), f32(0), f32(1)).astype(f32)
    return i0, i1, lam


def _lerp(t0, w0, t1, w1):
    """ATen UpSampleKernel Interpolate<>::eval as compiled for x86 FMA: fma(t0, w0, round(t1 * w1)).
    For the dyadic scales of the reference workloads (64->256, 224->256) every form is exact."""
    w0 = np.broadcast_to(w0, t0.shape).astype(np.float32)
    w1 = np.broadcast_to(w1, t0.shape).astype(np.float32)
    return _fma(t0, w0, t1 * w1)


def resize_bilinear_f32(frames_nchw_u8, rh, rw):
    """float32 bilinear resize (align_corners=False, no antialias) of uint8 NCHW frames, before rounding:
    bit-exact with torch.nn.functional.interpolate on CPU (tests/test_oracle_golden.py)."""
    x = np.asarray(frames_nchw_u8)
    n, c, h, w = x.shape
    f32 = np.float32
    y0, y1, ly = _src_index(f32(h) / f32(rh), np.arange(rh), h)
    x0, x1, lx = _src_index(f32(w) / f32(rw), np.arange(rw), w)
    hy, hx = (f32(1) - ly), (f32(1) - lx)
    xf = x.astype(f32)
    r0, r1 = xf[:, :, y0, :], xf[:, :, y1, :]
    top_ = _lerp(r0[..., x0], hx, r0[..., x1], lx)
    bot_ = _lerp(r1[..., x0], hx, r1[..., x1], lx)
    return _lerp(top_, hy[:, None], bot_, ly[:, None])


def _cubic_src_index(scale, dst, size):
    """ATen area_pixel_compute_source_index(align_corners=False, cubic=True): no clamp at zero; then
    guard_index_and_lambda (floor, index <= size-1, lambda clipped to [0, 1])."""
    f32 = np.float32
    d = dst.astype(f32) + f32(0.5)
    s = _fma(np.full_like(d, f32(scale)), d, np.full_like(d, f32(-0.5)))
    idx = np.minimum(np.floor(s).astype(np.int64), size - 1)
    lam = np.clip((s - idx.astype(f32)).astype(f32), f32(0), f32(1)).astype(f32)
    return idx, lam


def _cubic_coefficients(t):
    """ATen get_cubic_upsample_coefficients (A = -0.75) as compiled for x86 (probed bit-exact against
    F.interpolate on dyadic and non-dyadic sizes, see tests/test_oracle_golden.py):
      cubic_convolution1(x) = ((A+2)x - (A+3)) x x + 1        -> fma(1.25, x, -2.25), two rounded products, rounded + 1
      cubic_convolution2(x) = ((A x - 5A) x + 8A) x - 4A      -> fma(fma(-0.75, x, 3.75), x, -6), rounded product, + 3
    """
    f32 = np.float32

    def cc1(x):
        t1 = _fma(np.full_like(x, f32(1.25)), x, np.full_like(x, f32(-2.25)))
        return (((t1 * x).astype(f32) * x).astype(f32) + f32(1)).astype(f32)

    def cc2(x):
        t1 = _fma(np.full_like(x, f32(-0.75)), x, np.full_like(x, f32(3.75)))
        t2 = _fma(t1, x, np.full_like(x, f32(-6)))
        return ((t2 * x).astype(f32) + f32(3)).astype(f32)

    x1 = t.astype(f32)
    x2 = (f32(1) - x1).astype(f32)
    return [cc2((x1 + f32(1)).astype(f32)), cc1(x1), cc1(x2), cc2((x2 + f32(1)).astype(f32))]


def _cubic_sum(t, w):
    """ATen Interpolate<>::eval with four taps as compiled for x86 FMA:
    fma(t3, w3, fma(t2, w2, fma(t0, w0, round(t1 * w1))))."""
    w = [np.broadcast_to(wk, t[0].shape).astype(np.float32) for wk in w]
    o = _fma(t[0], w[0], (t[1] * w[1]).astype(np.float32))
    o = _fma(t[2], w[2], o)
    return _fma(t[3], w[3], o)


def resize_bicubic_f32(frames_nchw_u8, rh, rw):
    """float32 bicubic resize (align_corners=False, no antialias, A = -0.75) of uint8 NCHW frames before clamping /
    rounding: bit-exact with torch.nn.functional.interpolate(mode='bicubic') on CPU. Border taps are clamped
    (upsample_get_value_bounded)."""
    x = np.asarray(frames_nchw_u8)
    n, c, h, w = x.shape
    f32 = np.float32
    iy, ly = _cubic_src_index(f32(h) / f32(rh), np.arange(rh), h)
    ix, lx = _cubic_src_index(f32(w) / f32(rw), np.arange(rw), w)
    wy, wx = _cubic_coefficients(ly), _cubic_coefficients(lx)
    xf = x.astype(f32)
    rows = []
    for j in range(4):
        r = xf[:, :, np.clip(iy + j - 1, 0, h - 1), :]
        rows.append(_cubic_sum([r[..., np.clip(ix + k - 1, 0, w - 1)] for k in range(4)], wx))
    return _cubic_sum(rows, [wk[:, None] for wk in wy])


# ---- antialiased bicubic (CLIP transforms, src/embeddings.py:309-310: T.Resize(res, BICUBIC, antialias=True))
def _aa_cubic_filter(x):
    """ATen aa_filter for bicubic (a = -0.5, the PIL filter), as compiled for x86 (every a*b+c fused):
    |x| < 1: ((a+2)|x| - (a+3)) x^2 + 1;  |x| < 2: (((|x| - 5)|x| + 8)|x| - 4) a;  else 0."""
    f32 = np.float32
    x = np.abs(x).astype(f32)
    t = _fma(np.full_like(x, f32(1.5)), x, np.full_like(x, f32(-2.5)))
    near = _fma((t * x).astype(f32), x, np.full_like(x, f32(1)))
    u = _fma((x - f32(5)).astype(f32), x, np.full_like(x, f32(8)))
    u = _fma(u, x, np.full_like(x, f32(-4)))
    far = (u * f32(-0.5)).astype(f32)
    return np.where(x < 1, near, np.where(x < 2, far, f32(0))).astype(f32)


def _aa_weights(in_size, out_size):
    """ATen _compute_indices_min_size_weights_aa (UpSampleKernel.cpp) for one dimension: per output index the first
    input index and the normalised float32 weights. The C++ mixes float variables with double literals (`+ 0.5`,
    `1.0 / scale`): those sub-expressions are evaluated in double and rounded once — reproduced here, it changes a
    few weights by one ulp. Extracted weights are bit-identical with ATen's for all sizes tried (impulse inputs)."""
    f32, f64 = np.float32, np.float64
    scale = f32(in_size) / f32(out_size)
    support = f32(f32(2) * scale) if scale >= 1 else f32(2)
    invscale = f32(1.0 / f64(scale)) if scale >= 1 else f32(1)
    out = []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        xmin = max(int(f64(f32(center - support)) + 0.5), 0)
        xsize = min(int(f64(f32(center + support)) + 0.5), in_size) - xmin
        j = (np.arange(xsize) + xmin).astype(f32)
        w = _aa_cubic_filter(((f64(1) * (j - center).astype(f32) + 0.5) * f64(invscale)).astype(f32))
        total = f32(0)
        for v in w:
            total = f32(total + v)
        out.append((xmin, (w / total).astype(f32)))
    return out


def _aa_apply(x, weights, axis):
    """ATen interpolate_aa_single_dim along `axis`: out = x[xmin] * w[0], then `out += x[xmin + j] * w[j]`. As compiled
    for x86 the loop runs in groups of four iterations with a rounded product and a separate add, and the remaining
    (xsize - 1) mod 4 iterations as fused multiply-adds (probed: the only schedule that is bit-identical for 4-tap
    up-scaling and 12-tap down-scaling alike)."""
    f32 = np.float32
    x = np.moveaxis(x, axis, -1)
    out = np.empty(x.shape[:-1] + (len(weights),), f32)
    for i, (xmin, w) in enumerate(weights):
        o = (x[..., xmin] * w[0]).astype(f32)
        n = len(w) - 1
        grouped = n // 4 * 4
        for j in range(1, n + 1):
            if j <= grouped:
                o = (o + (x[..., xmin + j] * w[j]).astype(f32)).astype(f32)
            else:
                o = _fma(x[..., xmin + j], np.full(o.shape, w[j], f32), o)
        out[..., i] = o
    return np.moveaxis(out, -1, axis)


def resize_bicubic_aa_f32(frames_nchw_u8, rh, rw):
    """float32 antialiased bicubic resize of uint8 NCHW frames before clamping / rounding: the horizontal pass over
    all input rows into a float32 intermediate, then the vertical pass (ATen's separable order). Bit-exact with
    torch.nn.functional.interpolate(mode='bicubic', antialias=True, align_corners=False) on CPU."""
    x = np.asarray(frames_nchw_u8).astype(np.float32)
    h, w = x.shape[2:]
    if w != rw:
        x = _aa_apply(x, _aa_weights(w, rw), 3)
    if h != rh:
        x = _aa_apply(x, _aa_weights(h, rh), 2)
    return x


def resize_crop_u8(frames_nchw_u8, size=256, crop=224, interpolation="bilinear"):
    """Resize(256) + CenterCrop(224) on uint8 NCHW frames, bit-for-bit torchvision-0.10 semantics.

    Reference: src/embeddings.py:81-82. torchvision resizes uint8 tensors by casting to float32, bilinear
    interpolation with align_corners=False, torch.round (half to even) and a cast back to uint8
    (tv:transforms/_functional_tensor.py:462-472, 532-541).
    No antialiasing: torchvision 0.10 (the reference's pin, requirements.txt:4) has none. torchvision >= 0.17
    defaults to antialias=True, a different kernel; for the dyadic up-scalings the reference performs (64 -> 256,
    224 -> 256, 96 -> 256 ...) both are exact and identical, for non-dyadic ratios (e.g. 128 -> 341) they differ
    by one grey level on ~1e-5 of the pixels (exact .5 ties computed with different rounding).
    """
    x = np.asarray(frames_nchw_u8)
    rh, rw, top, left = resize_geometry(x.shape[2], x.shape[3], size, crop)
    if interpolation == "bicubic_aa":
        # CLIP: T.Resize(res, BICUBIC, antialias=True) (src/embeddings.py:310). torchvision returns the image untouched
        # when the short side already has the requested size (tv:transforms/functional.py:468-471)
        if (rh, rw) == x.shape[2:]:
            return np.ascontiguousarray(x[:, :, top:top + crop, left:left + crop])
        v = resize_bicubic_aa_f32(x, rh, rw)[:, :, top:top + crop, left:left + crop]
    elif interpolation == "bicubic":
        # T.Resize(256, interpolation=3) of the MAE encoders (src/embeddings.py:81): bicubic overshoots, torchvision
        # clamps to [0, 255] before the rounding cast (tv:transforms/_functional_tensor.py:469-470)
        v = resize_bicubic_f32(x, rh, rw)[:, :, top:top + crop, left:left + crop]
    else:
        v = resize_bilinear_f32(x, rh, rw)[:, :, top:top + crop, left:left + crop]
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)  # np.rint = half to even = torch.round


def normalize_lut(mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """ConvertImageDtype(float) + Normalize (src/embeddings.py:83-84) for every uint8 value: x.to(f32)/255
    (tv:transforms/_functional_tensor.py:97-99) then sub_(mean).div_(std) (:928) — three rounded fp32 ops."""
    f32 = np.float32
    u = np.arange(256, dtype=f32) / f32(255.0)
    m = np.asarray(mean, dtype=f32)[:, None]
    s = np.asarray(std, dtype=f32)[:, None]
    return ((u[None, :] - m) / s).astype(f32)  # (3, 256)


def transforms(frames_nchw_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD, interpolation="bilinear"):
    """Full reference `transforms` (src/embeddings.py:80-85) on (N,3,H,W) uint8 -> (N,3,224,224) float32."""
    u = resize_crop_u8(frames_nchw_u8, interpolation=interpolation)
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][u[:, c]] for c in range(3)], 1)


def split_frames(obs_nhwc):
    """(N,H,W,3n) -> (n*N,H,W,3) frame-major: main_bc_1.py:134, behavioral_cloning/save_embedded_obs.py:153."""
    n_frames = max(obs_nhwc.shape[3] // 3, 1)
    return np.concatenate(np.split(obs_nhwc, n_frames, axis=3), axis=0), n_frames


def regroup_frames(emb, n_frames):
    """(n*N, O) -> (N, O*n): main_bc_1.py:136, save_embedded_obs.py:155."""
    return np.concatenate(np.split(emb, n_frames, axis=0), axis=-1)


# ------------------------------------------------------------------------------------------------ ResNet-50 PVR
def _bn(x, sd, p):
    """eval-mode BatchNorm2d, eps 1e-5 (tv:models/resnet.py uses nn.BatchNorm2d defaults)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _bottleneck(x, sd, p, stride):
    """tv:models/resnet.py:143-166 (v1.5: stride on the 3x3 conv)."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if p + ".downsample.0.weight" in sd:
        x = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


def _basic_compress(x, sd, p):
    """tv:models/resnet.py:89-105 BasicBlock with the biased 3x3 downsample of src/vision_models/moco.py:34-50."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], padding=1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2")
    idn = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], sd[p + ".downsample.0.bias"], padding=1), sd,
              p + ".downsample.1")
    return F.relu(out + idn)


def resnet50_forward(sd, variant, x):
    """ResNet-50 forward (tv:models/resnet.py:266-282) with the reference's surgery:
    'conv5' = moco_conv5 / resnet_conv5 / resnet50 (src/vision_models/moco.py:6-26): fc -> identity, output (N, 2048);
    'l4'    = moco_conv4_compressed (moco.py:73-113): + BasicBlock(2048->42), avgpool blanked, output (N, 42*7*7);
    'l3'    = moco_conv3_compressed (moco.py:29-70): layer3 + BasicBlock(1024->11), layer4/avgpool blanked.
    x: (N,3,224,224) float32. Output flattened NCHW like src/embeddings.py:398 `.view(-1, out_size)`."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    l3 = "layer3.0." if variant == "l3" else "layer3."
    l4 = "layer4.0." if variant == "l4" else "layer4."
    for name, blocks, stride in (("layer1.", 3, 1), ("layer2.", 4, 2), (l3, 6, 2), (l4, 3, 2)):
        if name.startswith("layer4") and variant == "l3":
            break
        for b in range(blocks):
            x = _bottleneck(x, sd, f"{name}{b}", stride if b == 0 else 1)
    if variant == "l3":
        x = _basic_compress(x, sd, "layer3.1")
    elif variant == "l4":
        x = _basic_compress(x, sd, "layer4.1")
    else:
        x = F.adaptive_avg_pool2d(x, 1)
    return x.reshape(x.shape[0], -1)


RESNET_BASIC_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}


def resnet_basic_state(name, seed):
    """Deterministic random weights with torchvision's resnet18 / resnet34 key names (BasicBlock nets,
    tv:models/resnet.py:59-101, 266-282), fc omitted (the reference replaces it by Identity, src/embeddings.py:112-117)."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(key, co, ci, k):
        sd[key + ".weight"] = torch.from_numpy(
            (rng.standard_normal((co, ci, k, k), dtype=np.float32) * np.float32(math.sqrt(2.0 / (co * k * k)))))

    def bn(key, c, gain):
        sd[key + ".weight"] = torch.from_numpy(rng.uniform(gain * 0.75, gain * 1.25, c).astype(np.float32))
        sd[key + ".bias"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
        sd[key + ".running_mean"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
        sd[key + ".running_var"] = torch.from_numpy(rng.uniform(0.75, 1.25, c).astype(np.float32))
        sd[key + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", 64, 3, 7)
    bn("bn1", 64, 1.0)
    c_in = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), RESNET_BASIC_LAYERS[name])):
        for b in range(blocks):
            p = f"layer{li + 1}.{b}"
            conv(p + ".conv1", planes, c_in, 3)
            bn(p + ".bn1", planes, 1.0)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes, 0.5)
            if b == 0 and li > 0:
                conv(p + ".downsample.0", planes, c_in, 1)
                bn(p + ".downsample.1", planes, 0.7)
            c_in = planes
    return sd


def resnet_basic_forward(sd, name, x):
    """torchvision resnet18 / resnet34 forward with fc = Identity (src/embeddings.py:112-117): output (N, 512)."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, blocks in enumerate(RESNET_BASIC_LAYERS[name]):
        for b in range(blocks):
            p = f"layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            t = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn1"))
            t = _bn(F.conv2d(t, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2")
            if p + ".downsample.0.weight" in sd:
                x = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
            x = F.relu(t + x)
    return F.adaptive_avg_pool2d(x, 1).reshape(x.shape[0], -1)


def small_conv_forward(sd, x):
    """The 'random' PVR (src/embeddings.py:90-106): 5 x [Conv2d(3x3, s2, p1, bias) -> ELU], flattened NCHW."""
    for i in (0, 2, 4, 6, 8):
        x = F.elu(F.conv2d(x, sd[f"{i}.weight"], sd[f"{i}.bias"], stride=2, padding=1))
    return x.reshape(x.shape[0], -1)


def small_conv_embedding(sd, obs_nhwc_u8):
    x = torch.from_numpy(transforms(np.ascontiguousarray(np.transpose(obs_nhwc_u8, (0, 3, 1, 2)))))
    with torch.no_grad():
        return small_conv_forward({k: v.float() for k, v in sd.items()}, x).numpy()


UBER = {"345": ("l3", "l4", "conv5"), "35": ("l3", "conv5"), "34": ("l3", "l4"), "45": ("l4", "conv5")}


def embedding_forward(parts, obs_nhwc_u8):
    """EmbeddingNet.forward (src/embeddings.py:386-402) for one or several trunks (UberModel, :44-57).

    parts: list of (variant, state_dict); obs: (N,H,W,3) uint8 numpy. Returns float32 numpy (N, sum O)."""
    x = torch.from_numpy(transforms(np.ascontiguousarray(np.transpose(obs_nhwc_u8, (0, 3, 1, 2)))))
    with torch.no_grad():
        outs = [resnet50_forward(sd, variant, x) for variant, sd in parts]
    return torch.cat(outs, 1).numpy()


def embed_observations(parts, obs_nhwc_u8, batch_size=64):
    """The mini-batch embedding loop of main_bc_1.py:128-137 / save_embedded_obs.py:149-156:
    (N,H,W,3n) uint8 -> (N, O*n) float32."""
    out = []
    for i in range(0, obs_nhwc_u8.shape[0], batch_size):
        o, nf = split_frames(obs_nhwc_u8[i:i + batch_size])
        out.append(regroup_frames(embedding_forward(parts, o), nf))
    return np.concatenate(out)


# ------------------------------------------------------------------------------------------------ synthetic weights
RESNET50_LAYERS = (("layer1", 64, 3), ("layer2", 128, 4), ("layer3", 256, 6), ("layer4", 512, 3))


def resnet50_state(variant, seed):
    """Deterministic (numpy default_rng) random weights with torchvision's ResNet-50 key names and the reference's
    compressed-variant surgery; conv init is kaiming-normal fan_out like tv:models/resnet.py:208-210, BatchNorm
    statistics are non-trivial so BN folding is exercised."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, co, ci, k, bias=False):
        sd[name + ".weight"] = torch.from_numpy(
            (rng.standard_normal((co, ci, k, k), dtype=np.float32) * np.float32(math.sqrt(2.0 / (co * k * k)))))
        if bias:
            sd[name + ".bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, co).astype(np.float32))

    def bn(name, c, gain):
        sd[name + ".weight"] = torch.from_numpy(rng.uniform(gain * 0.75, gain * 1.25, c).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
        sd[name + ".running_mean"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
        sd[name + ".running_var"] = torch.from_numpy(rng.uniform(0.75, 1.25, c).astype(np.float32))
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", 64, 3, 7)
    bn("bn1", 64, 1.0)
    c_in = 64
    for name, planes, blocks in RESNET50_LAYERS:
        if name == "layer4" and variant == "l3":
            break
        base = name + (".0" if (name == "layer3" and variant == "l3") or (name == "layer4" and variant == "l4")
                       else "")
        for b in range(blocks):
            p = f"{base}.{b}"
            conv(p + ".conv1", planes, c_in, 1)
            bn(p + ".bn1", planes, 1.0)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes, 1.0)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4, 0.5)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, c_in, 1)
                bn(p + ".downsample.1", planes * 4, 0.7)
            c_in = planes * 4
    if variant in ("l3", "l4"):
        p, C, c = ("layer3.1", 1024, 11) if variant == "l3" else ("layer4.1", 2048, 42)
        conv(p + ".conv1", c, C, 3)
        bn(p + ".bn1", c, 1.0)
        conv(p + ".conv2", c, c, 3)
        bn(p + ".bn2", c, 1.0)
        conv(p + ".downsample.0", c, C, 3, bias=True)
        bn(p + ".downsample.1", c, 1.0)
    return sd


def structured_frames(n, h, w, ch, seed):
    """Synthetic frames that are not iid noise (SURVEY.md hard part 5): smooth gradients + random rectangles +
    per-frame brightness/contrast jitter + 5% noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h, w, ch), dtype=np.uint8)
    for i in range(n):
        img = np.empty((h, w, ch), dtype=np.float32)
        for c in range(ch):
            a, b, c0 = rng.uniform(-1, 1, 3)
            img[..., c] = 128 + 90 * (a * (xx / w - 0.5) + b * (yy / h - 0.5)) + 40 * c0
        for _ in range(int(rng.integers(3, 9))):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            y1 = min(h, y0 + int(rng.integers(4, h // 2 + 5)))
            x1 = min(w, x0 + int(rng.integers(4, w // 2 + 5)))
            img[y0:y1, x0:x1] = rng.uniform(0, 255, ch)
        img = (img - 128) * rng.uniform(0.6, 1.3) + 128 + rng.uniform(-30, 30)
        img += rng.normal(0, 0.05 * 255, img.shape)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def adversarial_frames(h, w):
    """Constant 0 / 255 frames, checkerboards and ramps whose bilinear samples land on .5 rounding ties."""
    yy, xx = np.mgrid[0:h, 0:w]
    frames = [np.zeros((h, w, 3), np.uint8), np.full((h, w, 3), 255, np.uint8)]
    chk = (((yy + xx) & 1) * 255).astype(np.uint8)
    frames.append(np.stack([chk, 255 - chk, chk], -1))
    ramp = ((xx * 4 + yy * 4) % 256).astype(np.uint8)  # differences of 4 with k/8 weights -> exact .5 ties
    frames.append(np.stack([ramp, ramp.T[:h, :w] if h == w else ramp, 255 - ramp], -1))
    odd = ((xx % 2) * 1 + (yy % 2) * 2 + 100).astype(np.uint8)
    frames.append(np.stack([odd, odd + 1, odd + 3], -1))
    return np.stack(frames)
