"""GPU bring-up cases: each compares one CUDA entry point with a plain PyTorch fp32 computation on the same inputs.
Run one case per process (`python tests/gpu_cases.py NAME`) so a device-side trap cannot poison the others;
`python tests/gpu_cases.py all` drives that. The pytest `-m gpu` suite imports the same functions.
"""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pvr_habitat_b200 import _lib  # noqa: E402
from pvr_habitat_b200 import program as prg  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def stats(name, got, ref, extra=None):
    d = (got.double() - ref.double()).abs()
    out = dict(case=name, rel_l2=rel_l2(got, ref), max_abs=float(d.max()), ref_absmax=float(ref.abs().max()),
               nan=int(torch.isnan(got.float()).sum()))
    if extra:
        out.update(extra)
    return out


# ------------------------------------------------------------------------------------------------ GEMM
def case_gemm(m=1000, n=256, k=512, relu=True, res=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16).cuda()
    n_pad = (n + 63) // 64 * 64
    b = torch.zeros(n_pad, k, dtype=torch.bfloat16)
    b[:n] = (torch.randn(n, k, generator=g) / k ** 0.5).to(torch.bfloat16)
    b = b.cuda()
    scale = torch.zeros(n_pad)
    bias = torch.zeros(n_pad)
    scale[:n] = torch.rand(n, generator=g) + 0.5
    bias[:n] = torch.randn(n, generator=g)
    scale, bias = scale.cuda(), bias.cuda()
    r = torch.randn(m, n, generator=g).to(torch.bfloat16).cuda() if res else None
    out = torch.full((m, n), float('nan'), dtype=torch.bfloat16, device='cuda')
    _lib.check(_lib.lib().pvr_gemm_bf16(a.data_ptr(), k, b.data_ptr(), k, out.data_ptr(), n, scale.data_ptr(),
                                        bias.data_ptr(), r.data_ptr() if res else None, n, m, n, n_pad, k,
                                        int(relu), _lib.current_stream_ptr()), "pvr_gemm_bf16")
    torch.cuda.synchronize()
    ref = a.float() @ b[:n].float().t() * scale[:n] + bias[:n]
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    return stats(f"gemm_{m}x{n}x{k}", out.float(), ref)


# ------------------------------------------------------------------------------------------------ single conv op
def run_conv_op(x_nhwc, w, scale, bias, stride, pad, relu, res_nhwc=None, stem=False, block_n=0):
    """x_nhwc: (N,H,W,C) bf16 cuda; w: (Co,Ci,R,S) fp32 cpu. Returns (N,P,Q,Co) bf16."""
    n, h, wd, c = x_nhwc.shape
    co, ci, r, s = w.shape
    prog = prg.Program()
    p = (h + 2 * pad - r) // stride + 1
    q = (wd + 2 * pad - s) // stride + 1
    n_pad = (co + 63) // 64 * 64
    if stem:
        x_nhwc = prg.expand_stem_input(x_nhwc)
        in_slot = prog.new_slot(h * (wd // 2) * 32)
        res = None
        out_slot = prog.conv(in_slot, (32, h, wd // 2), prg.pack_stem_weight(w, n_pad), 256, co, 7, 1, (2, 1), (-3, 0),
                             (p, q), scale, bias, co if relu else 0, block_n=block_n)
    else:
        in_slot = prog.new_slot(h * wd * c)
        res = None
        if res_nhwc is not None:
            res_slot = prog.new_slot(p * q * co)
            res = (res_slot, co, 0)
        out_slot = prog.conv(in_slot, (c, h, wd), prg.pack_conv_weight(w, n_pad), r * s * ci, co, r, s,
                             (stride, stride), (-pad, -pad), (p, q), scale, bias, co if relu else 0, res=res,
                             block_n=block_n)
    prog.emb_width = 1
    enc = prog.finish('cuda')
    enc.bind(n)
    ws = enc.workspace
    off = enc.slot_ptr(in_slot) - ws.data_ptr()
    ws[off:off + x_nhwc.numel() * 2].view(torch.bfloat16).copy_(x_nhwc.flatten())
    if res is not None:
        off = enc.slot_ptr(res[0]) - ws.data_ptr()
        ws[off:off + res_nhwc.numel() * 2].view(torch.bfloat16).copy_(res_nhwc.flatten())
    emb = torch.zeros(n, 1, device='cuda')
    enc.forward(emb, 1)
    torch.cuda.synchronize()
    return enc.slot_tensor(out_slot, (n, p, q, co))


def case_conv(n=3, h=14, wd=14, ci=128, co=128, r=3, stride=1, pad=1, relu=True, res=False, seed=0, block_n=0,
              name=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, ci, h, wd, generator=g).to(torch.bfloat16)
    w = (torch.randn(co, ci, r, r, generator=g) / (ci * r * r) ** 0.5).to(torch.bfloat16).float()
    scale = torch.rand(co, generator=g) + 0.5
    bias = torch.randn(co, generator=g)
    p = (h + 2 * pad - r) // stride + 1
    q = (wd + 2 * pad - r) // stride + 1
    rs = torch.randn(n, co, p, q, generator=g).to(torch.bfloat16) if res else None
    ref = F.conv2d(x.float().cuda(), w.cuda(), stride=stride, padding=pad) * scale.cuda()[None, :, None, None] \
        + bias.cuda()[None, :, None, None]
    if res:
        ref = ref + rs.float().cuda()
    if relu:
        ref = ref.relu()
    got = run_conv_op(x.permute(0, 2, 3, 1).contiguous().cuda(), w, scale, bias, stride, pad, relu,
                      rs.permute(0, 2, 3, 1).contiguous().cuda() if res else None, block_n=block_n)
    return stats(name or f"conv{r}x{r}_s{stride}_{ci}->{co}_{h}x{wd}_n{n}", got.float().permute(0, 3, 1, 2), ref)


def case_stem(n=2, hw=224, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, hw, hw, generator=g).to(torch.bfloat16)
    w = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(torch.bfloat16).float()
    scale = torch.rand(64, generator=g) + 0.5
    bias = torch.randn(64, generator=g)
    ref = (F.conv2d(x.float().cuda(), w.cuda(), stride=2, padding=3) * scale.cuda()[None, :, None, None]
           + bias.cuda()[None, :, None, None]).relu()
    x4 = torch.zeros(n, hw, hw, 4, dtype=torch.bfloat16)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    got = run_conv_op(x4.cuda(), w, scale, bias, 2, 3, True, stem=True)
    return stats(f"stem7x7_n{n}_{hw}", got.float().permute(0, 3, 1, 2), ref)


# ------------------------------------------------------------------------------------------------ preprocessing
def torch_transforms(x_nchw_u8, mean, std, size=256, crop=224):
    """Plain-torch statement of torchvision's Resize(256)->CenterCrop(224)->ConvertImageDtype->Normalize on uint8."""
    from pvr_habitat_b200.embeddings import resize_geometry
    n, c, h, w = x_nchw_u8.shape
    rh, rw, top, left = resize_geometry(h, w, size, crop)
    y = F.interpolate(x_nchw_u8.float(), size=(rh, rw), mode='bilinear', align_corners=False)
    y = torch.round(y).clamp(0, 255).to(torch.uint8)
    y = y[:, :, top:top + crop, left:left + crop]
    y = y.to(torch.float32) / 255.0
    m = torch.tensor(mean, dtype=torch.float32, device=y.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=y.device).view(1, 3, 1, 1)
    return (y - m) / s


def case_preprocess(n=5, h=224, w=224, nf=1, seed=0):
    from pvr_habitat_b200.embeddings import Transforms, IMAGENET_MEAN, IMAGENET_STD
    rng = np.random.default_rng(seed)
    obs = torch.from_numpy(rng.integers(0, 256, size=(n, h, w, 3 * nf), dtype=np.uint8))
    t = Transforms()
    out = torch.full((nf * n, 3, 224, 224), float('nan'), device='cuda')
    t.run(obs.cuda(), nf, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
    torch.cuda.synchronize()
    frames = torch.cat(torch.split(obs, 3, dim=3), 0).permute(0, 3, 1, 2).contiguous()  # frame-major
    ref = torch_transforms(frames, IMAGENET_MEAN, IMAGENET_STD)  # CPU fp32, the reference's arithmetic
    got = out.cpu()
    neq = int((got != ref).sum())
    st = stats(f"preprocess_{h}x{w}_nf{nf}_n{n}", got, ref, dict(mismatches=neq, total=ref.numel()))
    # bf16 NHWC4 variant, sample-major
    out4 = torch.zeros(n * nf, 224, 224, 4, dtype=torch.bfloat16, device='cuda')
    t.run(obs.cuda(), nf, out4.data_ptr(), _lib.PVR_FMT_NHWC4_BF16, True)
    torch.cuda.synchronize()
    ref4 = ref.view(nf, n, 3, 224, 224).permute(1, 0, 3, 4, 2).reshape(n * nf, 224, 224, 3).to(torch.bfloat16)
    st["nhwc4_mismatches"] = int((out4[..., :3].cpu() != ref4).sum()) + int((out4[..., 3] != 0).sum())
    return st


# ------------------------------------------------------------------------------------------------ whole network
def torch_resnet50_reference(sd, variant, x):
    """fp32 torchvision ResNet-50 with the reference's surgery (src/vision_models/moco.py), on x (N,3,224,224)."""
    import torchvision
    from torch import nn
    m = torchvision.models.resnet50()
    if variant == 'l3':
        ds = nn.Sequential(nn.Conv2d(1024, 11, 3, 1, 1), m._norm_layer(11))
        m.layer3 = nn.Sequential(m.layer3, torchvision.models.resnet.BasicBlock(1024, 11, stride=1,
                                                                                 norm_layer=m._norm_layer,
                                                                                 downsample=ds))
        m.layer4 = nn.Sequential()
        m.avgpool = nn.Sequential()
    elif variant == 'l4':
        ds = nn.Sequential(nn.Conv2d(2048, 42, 3, 1, 1), m._norm_layer(42))
        m.layer4 = nn.Sequential(m.layer4, torchvision.models.resnet.BasicBlock(2048, 42, stride=1,
                                                                                 norm_layer=m._norm_layer,
                                                                                 downsample=ds))
        m.avgpool = nn.Sequential()
    m.fc = nn.Sequential()
    missing = m.load_state_dict(sd, strict=True)
    m.eval().to(x.device)
    with torch.no_grad():
        return m(x).reshape(x.shape[0], -1)


def randomize_bn(model, seed):
    """Non-trivial BN statistics so the folded scale/bias path is exercised (random init has identity BN)."""
    g = torch.Generator().manual_seed(seed)
    for name, mod in model.named_modules():
        if hasattr(mod, 'running_var'):
            c = mod.weight.shape[0]
            last = name.endswith('bn3') or name.endswith('bn2') and 'layer' in name and '.1.bn2' in name
            mod.weight.data = (torch.rand(c, generator=g) * 0.5 + (0.25 if last else 0.75))
            mod.bias.data = torch.randn(c, generator=g) * 0.1
            mod.running_mean.data = torch.randn(c, generator=g) * 0.1
            mod.running_var.data = torch.rand(c, generator=g) * 0.5 + 0.75


def structured_frames(n, h, w, ch, seed):
    """Synthetic frames that are not iid noise (SURVEY hard part 5): gradients + rectangles + jitter + 5% noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h, w, ch), dtype=np.uint8)
    for i in range(n):
        img = np.empty((h, w, ch), dtype=np.float32)
        for c in range(ch):
            a, b, c0 = rng.uniform(-1, 1, 3)
            img[..., c] = 128 + 90 * (a * (xx / w - 0.5) + b * (yy / h - 0.5)) + 40 * c0
        for _ in range(rng.integers(3, 9)):
            y0, x0 = rng.integers(0, h), rng.integers(0, w)
            y1, x1 = min(h, y0 + rng.integers(4, h // 2 + 5)), min(w, x0 + rng.integers(4, w // 2 + 5))
            img[y0:y1, x0:x1] = rng.uniform(0, 255, ch)
        img = (img - 128) * rng.uniform(0.6, 1.3) + 128 + rng.uniform(-30, 30)
        img += rng.normal(0, 0.05 * 255, img.shape)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def case_resnet(variant='conv5', n=4, hw=64, seed=1):
    from pvr_habitat_b200.embeddings import EmbeddingNet, IMAGENET_MEAN, IMAGENET_STD
    from pvr_habitat_b200.vision_models.moco import allow_random_init
    name = {'conv5': 'moco_aug', 'l4': 'moco_aug_l4', 'l3': 'moco_aug_l3'}[variant]
    torch.manual_seed(seed)
    with allow_random_init():
        net = EmbeddingNet(name)
    randomize_bn(net.embedding, seed)
    net.invalidate()
    obs = torch.from_numpy(structured_frames(n, hw, hw, 3, seed))
    got = torch.from_numpy(np.atleast_2d(net(obs)))
    sd = {k: v.detach().cpu() for k, v in net.embedding.state_dict().items()}
    x = torch_transforms(obs.permute(0, 3, 1, 2).contiguous(), IMAGENET_MEAN, IMAGENET_STD)
    ref = torch_resnet50_reference(sd, variant, x.cuda()).cpu()
    cos = F.cosine_similarity(got.double(), ref.double(), dim=1)
    gc, rc = got - ref.mean(0, keepdim=True), ref - ref.mean(0, keepdim=True)
    return stats(f"resnet50_{variant}_n{n}_{hw}", got, ref,
                 dict(min_cos=float(cos.min()), rel_l2_centered=rel_l2(gc, rc), shape=list(got.shape)))


CASES = {
    "preprocess_224": lambda: case_preprocess(5, 224, 224, 1),
    "preprocess_64_nf2": lambda: case_preprocess(7, 64, 64, 2),
    "preprocess_96x128_nf3": lambda: case_preprocess(3, 96, 128, 3),
    "gemm_small": lambda: case_gemm(128, 64, 64, False, False),
    "gemm_k256": lambda: case_gemm(256, 64, 256, False, False),
    "gemm_tail": lambda: case_gemm(1000, 256, 512, True, True),
    "gemm_big": lambda: case_gemm(128 * 300 + 17, 512, 1024, True, True),
    "conv1x1_tiled": lambda: case_conv(3, 14, 14, 128, 256, 1, 1, 0, True, True),
    "conv3x3_s1": lambda: case_conv(3, 14, 14, 128, 128, 3, 1, 1, True, False),
    "conv3x3_s1_56": lambda: case_conv(2, 56, 56, 64, 64, 3, 1, 1, True, False),
    "conv3x3_s2": lambda: case_conv(3, 28, 28, 128, 128, 3, 2, 1, True, False),
    "conv1x1_s2": lambda: case_conv(3, 28, 28, 256, 512, 1, 2, 0, False, False),
    "conv3x3_7x7": lambda: case_conv(5, 7, 7, 512, 512, 3, 1, 1, True, True),
    "stem": lambda: case_stem(2, 224),
    "resnet_conv5_64": lambda: case_resnet('conv5', 4, 64),
    "resnet_conv5_224": lambda: case_resnet('conv5', 3, 224),
    "resnet_l4": lambda: case_resnet('l4', 3, 64),
    "resnet_l3": lambda: case_resnet('l3', 3, 64),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which != "all":
        print("RESULT " + json.dumps(CASES[which]()))
        return
    results = []
    names = sys.argv[2:] or list(CASES)
    for name in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                results.append(json.loads(line[-1][7:]))
                print(name, "->", line[-1][7:], flush=True)
            else:
                tail = (r.stdout + r.stderr)[-1500:]
                results.append(dict(case=name, error=tail))
                print(name, "-> FAILED rc", r.returncode, tail, flush=True)
        except subprocess.TimeoutExpired:
            results.append(dict(case=name, error="timeout"))
            print(name, "-> TIMEOUT", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_cases.json", "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
