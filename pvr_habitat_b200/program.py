"""Host-side compiler from a (torchvision-named) ResNet-50 state_dict to the flat op program executed by
libpvr_b200 (include/pvr_b200.h: pvr_op / pvr_encoder_*).

What the program computes is what the reference builds in src/vision_models/moco.py:6-113 out of
torchvision's ResNet (torchvision/models/resnet.py:108-166 Bottleneck, :59-105 BasicBlock, :266-282 forward):
eval-mode BatchNorm is folded into a per-channel fp32 scale/bias applied in the GEMM epilogue, activations are NHWC
bf16, weights are bf16 (C_out, K) with K ordered (tap_row, tap_col, channel).
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import pvr_op, pvr_slot

BN_EPS = 1e-5


def _round_up(x, m):
    return (x + m - 1) // m * m


def fold_bn(sd, prefix, conv_bias=None):
    """eval-mode BN(x) = (x - mean) / sqrt(var + eps) * gamma + beta  ->  x * scale + bias (fp32)."""
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(v + BN_EPS)
    bias = b - m * scale
    if conv_bias is not None:
        bias = bias + conv_bias.double() * scale
    return scale.float(), bias.float()


def pack_conv_weight(w, n_pad):
    """(C_out, C_in, R, S) fp32 -> bf16 (n_pad, R*S*C_in), K ordered (r, s, c)."""
    co, ci, r, s = w.shape
    k = r * s * ci
    out = torch.zeros(n_pad, k, dtype=torch.bfloat16)
    out[:co] = w.permute(0, 2, 3, 1).reshape(co, k).to(torch.bfloat16)
    return out


def pack_stem_weight(w, n_pad):
    """7x7 stride-2 stem over the W-expanded input (PVR_FMT_STEM_BF16): per output column q the input row holds the
    8 columns 2q-3 .. 2q+4 x 4 channels, so the filter is 7 row taps of 32 values.

    K index = r*32 + j*4 + c with j = filter column (j = 7 and c = 3 carry zero weights); 7 taps padded to 8
    (k_pad = 256).
    """
    co, ci, r, s = w.shape
    assert (ci, r, s) == (3, 7, 7)
    out = torch.zeros(n_pad, 8, 8, 4, dtype=torch.float32)  # (co, r, j, c)
    out[:co, :7, :7, :3] = w.permute(0, 2, 3, 1)
    return out.reshape(n_pad, 256).to(torch.bfloat16)


def expand_stem_input(x4):
    """(N, H, W, 4) -> (N, H, W/2, 32): the PVR_FMT_STEM_BF16 layout, in torch (tests / NCHW entry point only)."""
    n, h, w, c = x4.shape
    pad = torch.zeros(n, h, w + 8, c, dtype=x4.dtype, device=x4.device)
    pad[:, :, 4:4 + w] = x4
    cols = [pad[:, :, 1 + e:1 + e + w:2] for e in range(8)]  # entry 2q+1+e = column 2q-3+e
    return torch.stack(cols, 3).reshape(n, h, w // 2, 8 * c)


class Program:
    """Accumulates ops/slots; `finish(device)` uploads the packed weights and creates the pvr_encoder."""

    def __init__(self):
        self.ops = []          # dicts of pvr_op fields + host tensors
        self.slot_elems = []   # per-slot bf16 elements per image
        self.free = []
        self.deferred = []     # slots handed back once the NEXT conv has chosen its output slot
        self.emb_width = 0

    # ---- slots
    def new_slot(self, elems):
        self.slot_elems.append(int(elems))
        return len(self.slot_elems) - 1

    def alloc(self, elems):
        if self.free:
            s = self.free.pop()
            self.slot_elems[s] = max(self.slot_elems[s], int(elems))
            return s
        return self.new_slot(elems)

    def release(self, s):
        if s not in self.free:
            self.free.append(s)

    def release_after_next_conv(self, s):
        """Keep `s` out of the free list until the next conv op has been planned: lets libpvr_b200 fuse that conv
        into the kernel that still reads `s` (conv_b2b.cu) without its output aliasing a live input."""
        self.deferred.append(s)

    # ---- ops
    def conv(self, in_slot, in_chw, w_packed, k_pad, c_out, r, s, stride, lower, out_hw, scale, bias, relu_n,
             in_pitch=None, res=None, out_slot=None, out_pitch=None, out_coff=0, block_n=0, flops=None, act=0,
             in2=None, flags=0):
        c_in, h_in, w_in = in_chw
        n_pad = w_packed.shape[0]
        p, q = out_hw
        if out_pitch is None:
            out_pitch = _round_up(c_out, 8)
        if out_slot is None:
            out_slot = self.alloc(p * q * out_pitch)
        sc = torch.zeros(n_pad, dtype=torch.float32)
        bi = torch.zeros(n_pad, dtype=torch.float32)
        sc[:c_out] = scale
        bi[:c_out] = bias
        op = dict(kind=_lib.PVR_OP_CONV, in_slot=in_slot, out_slot=out_slot, res_slot=-1, c_in=c_in, h_in=h_in,
                  w_in=w_in, in_pitch=in_pitch if in_pitch is not None else c_in, c_out=c_out, h_out=p, w_out=q,
                  out_pitch=out_pitch, res_pitch=0, out_coff=out_coff, res_coff=0, r=r, s=s,
                  stride_h=stride[0], stride_w=stride[1], lower_h=lower[0], lower_w=lower[1], relu_n=relu_n,
                  block_n=block_n, k_pad=k_pad, n_pad=n_pad, emb_offset=0, act=act, flags=flags,
                  _weight=w_packed.contiguous(), _scale=sc, _bias=bi,
                  flops_per_image=int(flops if flops is not None else 2 * p * q * c_out * r * s * c_in))
        if res is not None:
            op.update(res_slot=res[0], res_pitch=res[1], res_coff=res[2])
        if in2 is not None:  # (slot, (c, h, w), pitch, stride): second 1x1 input, see include/pvr_b200.h
            op.update(in2_slot=in2[0], in2_c=in2[1][0], in2_h=in2[1][1], in2_w=in2[1][2], in2_pitch=in2[2],
                      in2_stride=in2[3])
        self.ops.append(op)
        for s_ in self.deferred:
            self.release(s_)
        self.deferred = []
        return out_slot

    def maxpool(self, in_slot, c, h, w, flags=0):
        p, q = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        out_slot = self.alloc(p * q * c * (2 if flags & _lib.PVR_OP_FP32 else 1))
        self.ops.append(dict(kind=_lib.PVR_OP_MAXPOOL, in_slot=in_slot, out_slot=out_slot, res_slot=-1, c_in=c,
                             h_in=h, w_in=w, in_pitch=c, c_out=c, h_out=p, w_out=q, out_pitch=c, flags=flags))
        return out_slot, p, q

    def avgpool2(self, in_slot, c, h, w, flags=0):
        """nn.AvgPool2d(2) on an NHWC slot (CLIP's ModifiedResNet); returns (slot, h // 2, w // 2)."""
        p, q = h // 2, w // 2
        out_slot = self.alloc(p * q * c * (2 if flags & _lib.PVR_OP_FP32 else 1))
        self.ops.append(dict(kind=_lib.PVR_OP_AVGPOOL2, in_slot=in_slot, out_slot=out_slot, res_slot=-1, c_in=c,
                             h_in=h, w_in=w, in_pitch=c, c_out=c, h_out=p, w_out=q, out_pitch=c, flags=flags))
        return out_slot, p, q

    def flatten(self, in_slot, c, h, w, pitch, emb_offset, flags=0):
        self.ops.append(dict(kind=_lib.PVR_OP_FLATTEN, in_slot=in_slot, out_slot=-1, res_slot=-1, c_in=c, h_in=h,
                             w_in=w, in_pitch=pitch, c_out=c, emb_offset=emb_offset, flags=flags))

    def avgpool(self, in_slot, c, h, w, emb_offset, flags=0):
        self.ops.append(dict(kind=_lib.PVR_OP_AVGPOOL, in_slot=in_slot, out_slot=-1, res_slot=-1, c_in=c, h_in=h,
                             w_in=w, in_pitch=c, c_out=c, emb_offset=emb_offset, flags=flags))

    def head_tail(self, in_slot, pitch, c, h, w, aux, emb_offset, taps=False, flags=0):
        """taps: the input slot holds float32 per-tap partial sums (9 x 2c per pixel, pitch in floats), see pvr_b200.h"""
        self.ops.append(dict(kind=_lib.PVR_OP_HEAD, in_slot=in_slot, out_slot=-1, res_slot=-1, c_in=2 * c, h_in=h,
                             w_in=w, in_pitch=pitch, c_out=c, emb_offset=emb_offset, act=1 if taps else 0,
                             flags=flags, _aux=aux.contiguous()))

    # ---- finalise
    def finish(self, device):
        return Encoder(self, device)


class Encoder:
    """Owns the device copies of the packed weights, the pvr_encoder handle and the bound workspace."""

    def __init__(self, prog, device):
        self.lib = _lib.lib()
        self.device = torch.device(device)
        self.emb_width = prog.emb_width
        self._keep = []
        ops = (pvr_op * len(prog.ops))()
        for i, d in enumerate(prog.ops):
            o = ops[i]
            for k, v in d.items():
                if not k.startswith("_") and k != "flops_per_image":
                    setattr(o, k, int(v))
            for key, field in (("_weight", "weight"), ("_scale", "scale"), ("_bias", "bias"), ("_aux", "aux")):
                if key in d:
                    t = d[key].to(self.device)
                    self._keep.append(t)
                    setattr(o, field, t.data_ptr())
        slots = (pvr_slot * len(prog.slot_elems))()
        for i, e in enumerate(prog.slot_elems):
            slots[i].elems_per_image = e
        handle = ctypes.c_void_p()
        _lib.check(self.lib.pvr_encoder_create(ops, len(prog.ops), slots, len(prog.slot_elems), prog.emb_width,
                                               ctypes.byref(handle)), "pvr_encoder_create")
        self.handle = handle
        self.n_ops = len(prog.ops)
        self.n_images = 0
        self.workspace = None
        self.slot0 = None
        self.op_meta = [{k: v for k, v in d.items() if not k.startswith("_")} for d in prog.ops]
        # device tensors of every op, by op index: lets a training loop overwrite packed weights / biases in place
        self.op_tensors, k = [], 0
        for d in prog.ops:
            ent = {}
            for key in ("_weight", "_scale", "_bias", "_aux"):
                if key in d:
                    ent[key[1:]] = self._keep[k]
                    k += 1
            self.op_tensors.append(ent)

    def bind(self, n_images):
        if n_images == self.n_images:
            return
        need = self.lib.pvr_encoder_workspace_bytes(self.handle, n_images)
        if need < 0:
            raise _lib.PvrError("pvr_encoder_workspace_bytes failed")
        if self.workspace is None or self.workspace.numel() < need + 1024:
            self.workspace = None
            self.workspace = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        slot0 = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pvr_encoder_bind(self.handle, n_images, base, need, ctypes.byref(slot0)),
                       "pvr_encoder_bind")
        self.slot0 = slot0.value
        self.n_images = n_images

    def forward(self, emb, emb_ld=None):
        """emb: float32 CUDA tensor with at least n_images rows of emb_ld floats."""
        if emb_ld is None:
            emb_ld = emb.stride(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pvr_encoder_forward(self.handle, emb.data_ptr(), emb_ld, _lib.current_stream_ptr()),
                       "pvr_encoder_forward")

    def forward_timed(self, emb, emb_ld=None):
        """Like forward, returns the per-op device time in ms (synchronises)."""
        if emb_ld is None:
            emb_ld = emb.stride(0)
        ms = (ctypes.c_float * self.n_ops)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pvr_encoder_forward_timed(self.handle, emb.data_ptr(), emb_ld,
                                                          _lib.current_stream_ptr(), ms), "pvr_encoder_forward_timed")
        return list(ms)

    def slot_ptr(self, slot):
        return self.lib.pvr_encoder_slot_ptr(self.handle, slot)

    def slot_tensor(self, slot, shape):
        """View a slot as a bf16 tensor (tests / per-layer parity). Copies out of the workspace."""
        ptr = self.slot_ptr(slot)
        off = ptr - self.workspace.data_ptr()
        n = math.prod(shape)
        return self.workspace[off:off + 2 * n].view(torch.bfloat16).view(*shape).clone()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pvr_encoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ ResNet-50
RESNET50_LAYERS = (("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 6, 2), ("layer4", 512, 3, 2))


def _conv_bn(prog, sd, conv_key, bn_key, in_slot, in_chw, stride, pad, relu, res=None, block_n=0):
    w = sd[conv_key + ".weight"].float()
    co, ci, r, s = w.shape
    assert ci == in_chw[0], (conv_key, ci, in_chw)
    scale, bias = fold_bn(sd, bn_key, sd.get(conv_key + ".bias"))
    n_pad = _round_up(co, 64)
    h, wd = in_chw[1], in_chw[2]
    p = (h + 2 * pad - r) // stride + 1
    q = (wd + 2 * pad - s) // stride + 1
    out = prog.conv(in_slot, in_chw, pack_conv_weight(w, n_pad), r * s * ci, co, r, s, (stride, stride),
                    (-pad, -pad), (p, q), scale, bias, co if relu else 0, res=res, block_n=block_n)
    return out, (co, p, q)


def _bottleneck(prog, sd, prefix, x_slot, x_chw, stride, has_ds, stride_in_1x1=False):
    """torchvision Bottleneck (resnet.py:143-166), stride on the 3x3 conv (v1.5); `stride_in_1x1`: on conv1 instead
    (detectron2's BottleneckBlock as src/vision_models/maskrcnn.py:52-56 configures it, the MSRA / Caffe layout).

    Blocks with a projection shortcut compute `relu(bn3(conv3(t2)) + bn_d(conv_d(x)))` as ONE GEMM over the
    concatenated K = [t2 channels | x channels]: both BN scales are folded into the weight rows, the bias is b3 + b_d.
    The shortcut tensor is neither written nor re-read (layer1.0: 411 MB + 411 MB per 256 frames)."""
    st1, st2 = (stride, 1) if stride_in_1x1 else (1, stride)
    t1, s1 = _conv_bn(prog, sd, prefix + ".conv1", prefix + ".bn1", x_slot, x_chw, st1, 0, True)
    t2, s2 = _conv_bn(prog, sd, prefix + ".conv2", prefix + ".bn2", t1, s1, st2, 1, True)
    prog.release(t1)
    if has_ds and x_chw[0] % 64 == 0 and s2[0] % 64 == 0:
        w3 = sd[prefix + ".conv3.weight"].float()
        wd = sd[prefix + ".downsample.0.weight"].float()
        co = w3.shape[0]
        sc3, b3 = fold_bn(sd, prefix + ".bn3", sd.get(prefix + ".conv3.bias"))
        scd, bd = fold_bn(sd, prefix + ".downsample.1", sd.get(prefix + ".downsample.0.bias"))
        wcat = torch.cat([w3.reshape(co, -1) * sc3[:, None], wd.reshape(co, -1) * scd[:, None]], 1)
        n_pad = _round_up(co, 64)
        k = wcat.shape[1]
        packed = torch.zeros(n_pad, k, dtype=torch.bfloat16)
        packed[:co] = wcat.to(torch.bfloat16)
        c2, p, q = s2
        y = prog.conv(t2, s2, packed, k, co, 1, 1, (1, 1), (0, 0), (p, q), torch.ones(co), b3 + bd, co,
                      in2=(x_slot, x_chw, x_chw[0], stride),
                      flops=2 * p * q * co * (c2 + x_chw[0]))
        prog.release_after_next_conv(t2)      # the next block's conv1 may run inside this kernel (conv_b2b.cu)
        prog.release_after_next_conv(x_slot)
        return y, (co, p, q)
    if has_ds:
        idn, sidn = _conv_bn(prog, sd, prefix + ".downsample.0", prefix + ".downsample.1", x_slot, x_chw, stride, 0,
                             False)
        prog.release(x_slot)
    else:
        idn, sidn = x_slot, x_chw
    # the deferred releases happen BEFORE this conv3 is planned for slots deferred earlier, and after the next block's
    # conv1 for t2 / idn: that conv1 may run inside this conv3's kernel (conv_b2b.cu)
    y, sy = _conv_bn(prog, sd, prefix + ".conv3", prefix + ".bn3", t2, s2, 1, 0, True, res=(idn, sidn[0], 0))
    prog.release_after_next_conv(t2)
    prog.release_after_next_conv(idn)
    return y, sy


def _compress_head(prog, sd, prefix, x_slot, x_chw, emb_offset):
    """BasicBlock(C, c, downsample=Sequential(Conv2d(C, c, 3, padding=1, bias=True), BN)) of moco.py:34-50/:78-94.

    conv1 (-> bn1 -> ReLU) and the biased downsample conv (-> BN) read the same input and are both 3x3 convolutions
    over C = 1024 / 2048 channels with only 2c = 22 / 84 outputs. As an implicit GEMM that is nine reads of the input
    for a 32- / 96-wide tile; instead ONE 1x1 GEMM computes the nine per-tap partial sums of every pixel,
    Z[q, tap*2c + j] = W_tap[j] . x[q] (N = 18c -> 256 / 768, float32 output), and the head kernel forms
    sum_tap Z[p + tap - 1, tap] in fp32, applies bn1 / bn_d (+ ReLU on conv1's half), then conv2 -> bn2 -> += identity ->
    ReLU and writes the NCHW-flattened float32 embedding columns.
    """
    w1 = sd[prefix + ".conv1.weight"].float()
    wd = sd[prefix + ".downsample.0.weight"].float()
    c = w1.shape[0]
    C, h, w = x_chw
    s1, b1 = fold_bn(sd, prefix + ".bn1")
    sdn, bdn = fold_bn(sd, prefix + ".downsample.1", sd[prefix + ".downsample.0.bias"])
    wcat = torch.cat([w1, wd], 0)                                  # (2c, C, 3, 3)
    wz = wcat.permute(2, 3, 0, 1).reshape(9 * 2 * c, C)            # row = tap*2c + j
    n_pad = _round_up(9 * 2 * c, 64)
    packed = torch.zeros(n_pad, C, dtype=torch.bfloat16)
    packed[:9 * 2 * c] = wz.to(torch.bfloat16)
    zslot = prog.alloc(h * w * n_pad * 2)                          # float32 values: two bf16 elements each
    prog.conv(x_slot, x_chw, packed, C, n_pad, 1, 1, (1, 1), (0, 0), (h, w), torch.ones(n_pad), torch.zeros(n_pad), 0,
              out_slot=zslot, out_pitch=n_pad, flags=_lib.PVR_CONV_OUT_F32, flops=2 * h * w * 2 * c * 9 * C)
    w2 = sd[prefix + ".conv2.weight"].float()  # (c, c, 3, 3) -> (co, r, s, ci)
    s2, b2 = fold_bn(sd, prefix + ".bn2")
    cp = _round_up(c, 4)
    w2t = torch.zeros(3, 3, c, cp)
    w2t[..., :c] = w2.permute(2, 3, 1, 0)                          # (r, s, ci, co): float4 = 4 output channels
    aux = torch.cat([w2t.reshape(-1), s2, b2, s1, sdn, b1, bdn, torch.zeros(8)]).float()
    aux = aux[:(aux.numel() // 4) * 4]
    prog.head_tail(zslot, n_pad, c, h, w, aux, emb_offset, taps=True)
    prog.release(zslot)
    return c * h * w


def add_resnet50(prog, sd, variant, in_slot, emb_offset, hw=224, compact_stem=False, stride_in_1x1=False):
    """Append one ResNet-50 trunk reading the W-expanded bf16 frames (PVR_FMT_STEM_BF16) in `in_slot`.

    variant: 'conv5' (moco_conv5 / resnet50: avg-pooled 2048), 'l4' (moco_conv4_compressed: 42*7*7 = 2058),
             'l3' (moco_conv3_compressed: 11*14*14 = 2156). Returns the number of embedding columns written.
    """
    pre = {"conv5": ("layer3.", "layer4."), "l4": ("layer3.", "layer4.0."), "l3": ("layer3.0.", None)}[variant]
    # stem: 7x7/2 as a 7x1-tap conv over the W-expanded input (see pack_stem_weight)
    scale, bias = fold_bn(sd, "bn1")
    p = (hw + 6 - 7) // 2 + 1
    # compact_stem: `in_slot` holds padded NHWC4 rows (PVR_FMT_STEM_PAD_BF16); the 8-column windows are formed by the
    # stem's tensor map (pixel pitch 8 elements = 2 columns) instead of being materialised (pixel pitch 32)
    stem = prog.conv(in_slot, (32, hw, hw // 2), pack_stem_weight(sd["conv1.weight"].float(), 64), 256, 64, 7, 1,
                     (2, 1), (-3, 0), (p, p), scale, bias, 64, flops=2 * p * p * 64 * 147,
                     in_pitch=8 if compact_stem else None)
    x, h, w = prog.maxpool(stem, 64, p, p)
    prog.release(stem)
    chw = (64, h, w)
    for name, planes, blocks, stride in RESNET50_LAYERS:
        if name == "layer4" and pre[1] is None:
            break
        key = name + "."
        if name == "layer3":
            key = pre[0]
        elif name == "layer4":
            key = pre[1]
        for b in range(blocks):
            x, chw = _bottleneck(prog, sd, f"{key}{b}", x, chw, stride if b == 0 else 1, b == 0, stride_in_1x1)
    if variant == "conv5":
        prog.avgpool(x, chw[0], chw[1], chw[2], emb_offset)
        prog.release(x)
        return chw[0]
    head_prefix = "layer4.1" if variant == "l4" else "layer3.1"
    n = _compress_head(prog, sd, head_prefix, x, chw, emb_offset)
    prog.release(x)
    return n


# ------------------------------------------------------------------------------------------------ CLIP ModifiedResNet
def _clip_down_block(prog, sd, prefix, x_slot, x_chw):
    """First block of layer2-4 of CLIP's ModifiedResNet (openai/CLIP clip/model.py Bottleneck with stride 2): every conv
    has stride 1; `avgpool(2)` follows conv2 and precedes the shortcut's 1x1 conv. conv3 and the shortcut conv then run
    as ONE GEMM over K = [pooled t2 | pooled x] like torchvision's projection blocks (_bottleneck)."""
    t1, s1 = _conv_bn(prog, sd, prefix + ".conv1", prefix + ".bn1", x_slot, x_chw, 1, 0, True)
    t2, s2 = _conv_bn(prog, sd, prefix + ".conv2", prefix + ".bn2", t1, s1, 1, 1, True)
    prog.release(t1)
    t2p, p, q = prog.avgpool2(t2, s2[0], s2[1], s2[2])
    prog.release(t2)
    xp, _, _ = prog.avgpool2(x_slot, x_chw[0], x_chw[1], x_chw[2])
    prog.release(x_slot)
    w3 = sd[prefix + ".conv3.weight"].float()
    wd = sd[prefix + ".downsample.0.weight"].float()
    co = w3.shape[0]
    sc3, b3 = fold_bn(sd, prefix + ".bn3")
    scd, bd = fold_bn(sd, prefix + ".downsample.1")
    wcat = torch.cat([w3.reshape(co, -1) * sc3[:, None], wd.reshape(co, -1) * scd[:, None]], 1)
    k = wcat.shape[1]
    packed = torch.zeros(_round_up(co, 64), k, dtype=torch.bfloat16)
    packed[:co] = wcat.to(torch.bfloat16)
    y = prog.conv(t2p, (s2[0], p, q), packed, k, co, 1, 1, (1, 1), (0, 0), (p, q), torch.ones(co), b3 + bd, co,
                  in2=(xp, (x_chw[0], p, q), x_chw[0], 1), flops=2 * p * q * co * (s2[0] + x_chw[0]))
    prog.release_after_next_conv(t2p)
    prog.release_after_next_conv(xp)
    return y, (co, p, q)


def add_clip_resnet(prog, sd, in_slot, hw=224, layers=(3, 4, 6, 3)):
    """The convolutional trunk of CLIP's ModifiedResNet (`clip.load("RN50").visual` without its attention pool):
    3-conv stem (3 -> 32 /2, 32 -> 32, 32 -> 64, each + BN + ReLU) -> avgpool(2) -> layer1..4 (Bottlenecks with the
    stride replaced by average pooling). `in_slot` holds NHWC4 bf16 frames. Returns (slot, (2048, hw/32, hw/32)) of the
    NHWC bf16 feature map the attention pool reads; nothing is written to the embedding row."""
    h = hw
    p = (h + 2 - 3) // 2 + 1
    sc, bi = fold_bn(sd, "bn1")
    x = prog.conv(in_slot, (8, h, h // 2), pack_first_small_conv(sd["conv1.weight"].float(), 32), 64, 32, 3, 2, (2, 1),
                  (-1, -1), (p, p), sc, bi, 32, out_pitch=32, flops=2 * p * p * 32 * 27)
    for i, co in ((2, 32), (3, 64)):
        sc, bi = fold_bn(sd, f"bn{i}")
        y = prog.conv(x, (32, p, p), pack_small_conv(sd[f"conv{i}.weight"].float(), _round_up(co, 32)), 320, co, 3, 3,
                      (1, 1), (-1, -1), (p, p), sc, bi, co, out_pitch=co, flops=2 * p * p * co * 288)
        prog.release(x)
        x = y
    y, h, w = prog.avgpool2(x, 64, p, p)
    prog.release(x)
    x, chw = y, (64, h, w)
    for li, blocks in enumerate(layers):
        for b in range(blocks):
            prefix = f"layer{li + 1}.{b}"
            if b == 0 and li > 0:
                x, chw = _clip_down_block(prog, sd, prefix, x, chw)
            else:
                x, chw = _bottleneck(prog, sd, prefix, x, chw, 1, b == 0)
    return x, chw


def add_clip_resnet_f32(prog, sd, in_slot, hw=224, layers=(3, 4, 6, 3)):
    """fp32 counterpart of add_clip_resnet (float32 NHWC4 frames in, float32 NHWC feature map out)."""
    x, chw = in_slot, (4, hw, hw)
    for i, (stride, w) in enumerate(((2, _pad_rgb_weight(sd["conv1.weight"].float())), (1, sd["conv2.weight"].float()),
                                     (1, sd["conv3.weight"].float())), 1):
        sc, bi = fold_bn(sd, f"bn{i}")
        y, sy = _conv_f32(prog, w, sc, bi, x, chw, stride, 1, w.shape[0])
        if i > 1:
            prog.release(x)
        x, chw = y, sy
    y, h, w = prog.avgpool2(x, chw[0], chw[1], chw[2], flags=F32)
    prog.release(x)
    x, chw = y, (chw[0], h, w)
    for li, blocks in enumerate(layers):
        for b in range(blocks):
            px = f"layer{li + 1}.{b}"
            t1, s1 = _conv_bn_f32(prog, sd, px + ".conv1", px + ".bn1", x, chw, 1, 0, True)
            t2, s2 = _conv_bn_f32(prog, sd, px + ".conv2", px + ".bn2", t1, s1, 1, 1, True)
            prog.release(t1)
            if b == 0 and li > 0:
                t2p, h, w = prog.avgpool2(t2, s2[0], s2[1], s2[2], flags=F32)
                prog.release(t2)
                t2, s2 = t2p, (s2[0], h, w)
                xp, _, _ = prog.avgpool2(x, chw[0], chw[1], chw[2], flags=F32)
                prog.release(x)
                x, chw = xp, (chw[0], h, w)
            if b == 0:
                idn, sidn = _conv_bn_f32(prog, sd, px + ".downsample.0", px + ".downsample.1", x, chw, 1, 0, False)
                prog.release(x)
            else:
                idn, sidn = x, chw
            y, sy = _conv_bn_f32(prog, sd, px + ".conv3", px + ".bn3", t2, s2, 1, 0, True, res=(idn, sidn[0], 0))
            prog.release(t2)
            prog.release(idn)
            x, chw = y, sy
    return x, chw


def add_resnet_basic(prog, sd, layers, in_slot, emb_offset, hw=224, compact_stem=False):
    """Append a BasicBlock ResNet (resnet18: layers (2,2,2,2); resnet34: (3,4,6,3); tv:models/resnet.py:59-101) with
    fc = Identity (src/embeddings.py:112-117), reading the W-expanded frames in `in_slot`. Writes 512 columns."""
    scale, bias = fold_bn(sd, "bn1")
    p = (hw + 6 - 7) // 2 + 1
    # compact_stem: `in_slot` holds padded NHWC4 rows (PVR_FMT_STEM_PAD_BF16); the 8-column windows are formed by the
    # stem's tensor map (pixel pitch 8 elements = 2 columns) instead of being materialised (pixel pitch 32)
    stem = prog.conv(in_slot, (32, hw, hw // 2), pack_stem_weight(sd["conv1.weight"].float(), 64), 256, 64, 7, 1,
                     (2, 1), (-3, 0), (p, p), scale, bias, 64, flops=2 * p * p * 64 * 147,
                     in_pitch=8 if compact_stem else None)
    x, h, w = prog.maxpool(stem, 64, p, p)
    prog.release(stem)
    chw = (64, h, w)
    for li, blocks in enumerate(layers):
        for b in range(blocks):
            pre = f"layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            t, st = _conv_bn(prog, sd, pre + ".conv1", pre + ".bn1", x, chw, stride, 1, True)
            if pre + ".downsample.0.weight" in sd:
                idn, sidn = _conv_bn(prog, sd, pre + ".downsample.0", pre + ".downsample.1", x, chw, stride, 0, False)
                prog.release(x)
            else:
                idn, sidn = x, chw
            y, sy = _conv_bn(prog, sd, pre + ".conv2", pre + ".bn2", t, st, 1, 1, True, res=(idn, sidn[0], 0))
            prog.release(t)
            prog.release(idn)
            x, chw = y, sy
    prog.avgpool(x, chw[0], chw[1], chw[2], emb_offset)
    prog.release(x)
    return chw[0]


# ------------------------------------------------------------------------------------------------ fp32 parity mode
# The same networks with float32 activations / weights / accumulation on the CUDA cores (csrc/conv_f32.cu), op flag
# PVR_OP_FP32: the north star's "relative L2 <= 1e-5 in the fp32 mode". Frames arrive as PVR_FMT_NHWC4_F32 (RGB + a
# zero fourth channel); slot sizes count bf16 elements, so every float is two of them. No fusion, no packing tricks:
# one op per reference module, K ordered (tap_row, tap_col, channel).
F32 = _lib.PVR_OP_FP32


def _conv_f32(prog, w, scale, bias, in_slot, in_chw, stride, pad, relu_n, res=None, act=0):
    """w: (C_out, C_in, R, S) float32 with C_in == in_chw[0] (a multiple of 4). Returns (slot, (C_out, P, Q))."""
    co, ci, r, s = w.shape
    assert ci == in_chw[0] and ci % 4 == 0, (w.shape, in_chw)
    h, wd = in_chw[1], in_chw[2]
    p = (h + 2 * pad - r) // stride + 1
    q = (wd + 2 * pad - s) // stride + 1
    wk = w.permute(0, 2, 3, 1).reshape(co, r * s * ci).float().contiguous()
    out_slot = prog.alloc(p * q * co * 2)
    prog.conv(in_slot, in_chw, wk, r * s * ci, co, r, s, (stride, stride), (-pad, -pad), (p, q), scale, bias, relu_n,
              res=res, out_slot=out_slot, out_pitch=co, act=act, flags=F32)
    return out_slot, (co, p, q)


def _conv_bn_f32(prog, sd, conv_key, bn_key, in_slot, in_chw, stride, pad, relu, res=None):
    w = sd[conv_key + ".weight"].float()
    scale, bias = fold_bn(sd, bn_key, sd.get(conv_key + ".bias"))
    return _conv_f32(prog, w, scale, bias, in_slot, in_chw, stride, pad, w.shape[0] if relu else 0, res=res)


def _pad_rgb_weight(w):
    """(C_out, 3, R, S) -> (C_out, 4, R, S): the zero fourth channel of the NHWC4 frames."""
    co, ci, r, s = w.shape
    out = torch.zeros(co, 4, r, s)
    out[:, :ci] = w
    return out


def _stem_f32(prog, sd, in_slot, hw):
    scale, bias = fold_bn(sd, "bn1")
    stem, chw = _conv_f32(prog, _pad_rgb_weight(sd["conv1.weight"].float()), scale, bias, in_slot, (4, hw, hw), 2, 3,
                          64)
    x, h, w = prog.maxpool(stem, 64, chw[1], chw[2], flags=F32)
    prog.release(stem)
    return x, (64, h, w)


def add_resnet50_f32(prog, sd, variant, in_slot, emb_offset, hw=224, stride_in_1x1=False):
    """fp32 counterpart of add_resnet50 (same variants, same embedding columns)."""
    pre = {"conv5": ("layer3.", "layer4."), "l4": ("layer3.", "layer4.0."), "l3": ("layer3.0.", None)}[variant]
    x, chw = _stem_f32(prog, sd, in_slot, hw)
    for name, planes, blocks, stride in RESNET50_LAYERS:
        if name == "layer4" and pre[1] is None:
            break
        key = {"layer3": pre[0], "layer4": pre[1]}.get(name, name + ".")
        for b in range(blocks):
            px, st = f"{key}{b}", (stride if b == 0 else 1)
            st1, st2 = (st, 1) if stride_in_1x1 else (1, st)
            t1, s1 = _conv_bn_f32(prog, sd, px + ".conv1", px + ".bn1", x, chw, st1, 0, True)
            t2, s2 = _conv_bn_f32(prog, sd, px + ".conv2", px + ".bn2", t1, s1, st2, 1, True)
            prog.release(t1)
            if b == 0:
                idn, sidn = _conv_bn_f32(prog, sd, px + ".downsample.0", px + ".downsample.1", x, chw, st, 0, False)
                prog.release(x)
            else:
                idn, sidn = x, chw
            y, sy = _conv_bn_f32(prog, sd, px + ".conv3", px + ".bn3", t2, s2, 1, 0, True, res=(idn, sidn[0], 0))
            prog.release(t2)
            prog.release(idn)
            x, chw = y, sy
    if variant == "conv5":
        prog.avgpool(x, chw[0], chw[1], chw[2], emb_offset, flags=F32)
        prog.release(x)
        return chw[0]
    # compression head (moco.py:34-50 / :78-94): [conv1 -> bn1 -> ReLU | biased downsample conv -> BN] as one 3x3 conv
    # with 2c outputs, then conv2 -> bn2 -> += identity -> ReLU in the head kernel
    hp = "layer4.1" if variant == "l4" else "layer3.1"
    w1, wd = sd[hp + ".conv1.weight"].float(), sd[hp + ".downsample.0.weight"].float()
    c = w1.shape[0]
    s1, b1 = fold_bn(sd, hp + ".bn1")
    sdn, bdn = fold_bn(sd, hp + ".downsample.1", sd[hp + ".downsample.0.bias"])
    t, st = _conv_f32(prog, torch.cat([w1, wd], 0), torch.cat([s1, sdn]), torch.cat([b1, bdn]), x, chw, 1, 1, c)
    prog.release(x)
    s2, b2 = fold_bn(sd, hp + ".bn2")
    w2 = sd[hp + ".conv2.weight"].float().permute(0, 2, 3, 1).reshape(-1)  # (co, r, s, ci)
    prog.head_tail(t, 2 * c, c, chw[1], chw[2], torch.cat([w2, s2, b2]).float(), emb_offset, flags=F32)
    prog.release(t)
    return c * chw[1] * chw[2]


def add_resnet_basic_f32(prog, sd, layers, in_slot, emb_offset, hw=224):
    """fp32 counterpart of add_resnet_basic."""
    x, chw = _stem_f32(prog, sd, in_slot, hw)
    for li, blocks in enumerate(layers):
        for b in range(blocks):
            pre = f"layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            t, st = _conv_bn_f32(prog, sd, pre + ".conv1", pre + ".bn1", x, chw, stride, 1, True)
            if pre + ".downsample.0.weight" in sd:
                idn, sidn = _conv_bn_f32(prog, sd, pre + ".downsample.0", pre + ".downsample.1", x, chw, stride, 0,
                                         False)
                prog.release(x)
            else:
                idn, sidn = x, chw
            y, sy = _conv_bn_f32(prog, sd, pre + ".conv2", pre + ".bn2", t, st, 1, 1, True, res=(idn, sidn[0], 0))
            prog.release(t)
            prog.release(idn)
            x, chw = y, sy
    prog.avgpool(x, chw[0], chw[1], chw[2], emb_offset, flags=F32)
    prog.release(x)
    return chw[0]


def add_small_conv_f32(prog, sd, in_slot, emb_offset, hw=224):
    """fp32 counterpart of add_small_conv (src/embeddings.py:90-106)."""
    x, chw = in_slot, (4, hw, hw)
    for i in (0, 2, 4, 6, 8):
        w = sd[f"{i}.weight"].float()
        if i == 0:
            w = _pad_rgb_weight(w)
        y, chw = _conv_f32(prog, w, torch.ones(32), sd[f"{i}.bias"].float(), x, chw, 2, 1, 0, act=3)
        if x != in_slot:
            prog.release(x)
        x = y
    prog.flatten(x, 32, chw[1], chw[2], 32, emb_offset, flags=F32)
    prog.release(x)
    return 32 * chw[1] * chw[2]


# ------------------------------------------------------------------------------------------------ small-conv PVR
def pack_first_small_conv(w, n_pad):
    """3x3 stride-2 pad-1 conv over NHWC4 frames seen as pixel pairs (H, W/2, 8): output column q reads input columns
    2q-1 .. 2q+1 = pair q-1 (second pixel) and pair q (both pixels) -> 3 x 2 taps of 8 values, K padded to 64.
    K index = (r*2 + sp)*8 + e*4 + c with filter column j = 2*sp + e - 1."""
    co, ci, r, s = w.shape
    assert (ci, r, s) == (3, 3, 3)
    # (co, r (4: the last is padding), sp, e, c); tap = r*2 + sp. Works on any device (the finetuning path repacks
    # the weights on the GPU every step, without a host round trip).
    out = torch.zeros(n_pad, 4, 2, 2, 4, dtype=torch.float32, device=w.device)
    for (sp, e), j in (((0, 1), 0), ((1, 0), 1), ((1, 1), 2)):
        out[:co, :3, sp, e, :3] = w[:, :, :, j].permute(0, 2, 1)
    return out.reshape(n_pad, 64).to(torch.bfloat16)


def pack_small_conv(w, n_pad):
    """(32, 32, 3, 3) -> bf16 (n_pad, 320): K = (r, s, c) over 9 taps of 32 channels, padded to 10 taps."""
    co, ci, r, s = w.shape
    out = torch.zeros(n_pad, 10, ci, dtype=torch.float32, device=w.device)
    out[:co, :9] = w.permute(0, 2, 3, 1).reshape(co, 9, ci)
    return out.reshape(n_pad, 10 * ci).to(torch.bfloat16)


def add_small_conv(prog, sd, in_slot, emb_offset, hw=224, keep_activations=False):
    """The reference's 'random' PVR (src/embeddings.py:90-106): 5 x [Conv2d(3x3, stride 2, padding 1, bias) + ELU],
    3 -> 32 -> 32 -> 32 -> 32 -> 32 channels, output flattened NCHW. `in_slot` holds NHWC4 bf16 frames.
    Layer 1 uses 8-element pixel pairs (A_IM2COL8), layers 2-5 32-channel pixels (A_IM2COL32, 64-byte TMA rows);
    bias + ELU run in the GEMM epilogue. Returns the number of embedding columns."""
    ones = torch.ones(32)
    h = hw
    p = (h + 2 - 3) // 2 + 1
    x = prog.conv(in_slot, (8, h, h // 2), pack_first_small_conv(sd["0.weight"].float(), 32), 64, 32, 3, 2, (2, 1),
                  (-1, -1), (p, p), ones, sd["0.bias"].float(), 0, out_pitch=32, act=3, flops=2 * p * p * 32 * 27)
    kept = [x]
    h = p
    for i in (2, 4, 6, 8):
        p = (h + 2 - 3) // 2 + 1
        y = prog.conv(x, (32, h, h), pack_small_conv(sd[f"{i}.weight"].float(), 32), 320, 32, 3, 3, (2, 2), (-1, -1),
                      (p, p), ones, sd[f"{i}.bias"].float(), 0, out_pitch=32, act=3, flops=2 * p * p * 32 * 288)
        kept.append(y)
        if not keep_activations:  # training keeps every layer output for the backward pass
            prog.release(x)
        x, h = y, p
    prog.flatten(x, 32, h, h, 32, emb_offset)
    if not keep_activations:
        prog.release(x)
    prog.kept_slots = kept
    return 32 * h * h
