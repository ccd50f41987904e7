"""CPU emulation of the op program that pvr_habitat_b200.program emits (test infrastructure only).

It interprets the very fields libpvr_b200 consumes (packed bf16 weights, lower corner, traversal strides, pitches,
channel offsets), in fp32 torch, so host-side packing / slot planning is verified without a GPU.
"""
import torch
import torch.nn.functional as F

from pvr_habitat_b200 import _lib


def emulate(prog, frames_nhwc4, round_bf16=True, return_slots=False):
    """frames_nhwc4: slot-0 contents, (N, H, W/2, 32) W-expanded frames (program.expand_stem_input). Returns (N, emb_width);
    with `return_slots` also the final contents of every slot (programs that leave a feature map for a runner)."""
    n = frames_nhwc4.shape[0]
    slots = {0: frames_nhwc4.reshape(n, -1).float()}
    emb = torch.zeros(n, prog.emb_width)

    def rb(t):
        return t.to(torch.bfloat16).float() if round_bf16 else t

    slots[0] = rb(slots[0])
    for op in prog.ops:
        k = op["kind"]
        if k == _lib.PVR_OP_CONV:
            ci, hi, wi, pitch = op["c_in"], op["h_in"], op["w_in"], op["in_pitch"]
            x = slots[op["in_slot"]][:, :hi * wi * pitch].reshape(n, hi, wi, pitch)[..., :ci].permute(0, 3, 1, 2)
            r, s = op["r"], op["s"]
            w = op["_weight"].float()
            co, npad = op["c_out"], op["n_pad"]
            kk = r * s * ci
            wt = w[:, :kk].reshape(npad, r, s, ci).permute(0, 3, 1, 2)
            c2 = op.get("in2_c", 0)
            assert torch.all(w[:, kk + c2:] == 0), "padded K columns must carry zero weights"
            p, q = op["h_out"], op["w_out"]
            pl, pt = -op["lower_w"], -op["lower_h"]
            pr = (q - 1) * op["stride_w"] + s - wi - pl
            pb = (p - 1) * op["stride_h"] + r - hi - pt
            xp = F.pad(x, (pl, max(pr, 0), pt, max(pb, 0)))
            y = F.conv2d(xp, wt, stride=(op["stride_h"], op["stride_w"]))[:, :, :p, :q]
            if c2:  # second 1x1 input accumulated into the same GEMM (projection shortcut)
                h2, w2_, p2, st2 = op["in2_h"], op["in2_w"], op["in2_pitch"], op["in2_stride"]
                x2 = slots[op["in2_slot"]][:, :h2 * w2_ * p2].reshape(n, h2, w2_, p2)[..., :c2].permute(0, 3, 1, 2)
                wt2 = w[:, kk:kk + c2].reshape(npad, c2, 1, 1)
                y = y + F.conv2d(x2, wt2, stride=st2)[:, :, :p, :q]
            y = y * op["_scale"][None, :, None, None] + op["_bias"][None, :, None, None]
            y = y[:, :co]
            if op["res_slot"] >= 0:
                rp = op["res_pitch"]
                res = slots[op["res_slot"]][:, :p * q * rp].reshape(n, p, q, rp)
                y = y + res[..., op["res_coff"]:op["res_coff"] + co].permute(0, 3, 1, 2)
            rn = op["relu_n"]
            if rn > 0:
                y = torch.cat([y[:, :rn].relu(), y[:, rn:]], 1)
            if op.get("act", 0) == 3:
                y = F.elu(y)
            op_ = op["out_pitch"]
            buf = slots.get(op["out_slot"])
            need = p * q * op_
            if buf is None or buf.shape[1] < need:
                buf = torch.zeros(n, need)
            view = buf[:, :need].reshape(n, p, q, op_).clone()
            f32_out = op.get("flags", 0) & _lib.PVR_CONV_OUT_F32  # float32 slot: no bf16 rounding of the result
            view[..., op["out_coff"]:op["out_coff"] + co] = y.permute(0, 2, 3, 1) if f32_out else rb(y.permute(0, 2, 3, 1))
            nb = torch.zeros(n, max(need, buf.shape[1]))
            nb[:, :need] = view.reshape(n, -1)
            slots[op["out_slot"]] = nb
        elif k == _lib.PVR_OP_MAXPOOL:
            c, h, w = op["c_in"], op["h_in"], op["w_in"]
            x = slots[op["in_slot"]][:, :h * w * c].reshape(n, h, w, c).permute(0, 3, 1, 2)
            y = F.max_pool2d(x, 3, 2, 1)
            slots[op["out_slot"]] = y.permute(0, 2, 3, 1).reshape(n, -1).contiguous()
        elif k == _lib.PVR_OP_AVGPOOL2:
            c, h, w = op["c_in"], op["h_in"], op["w_in"]
            x = slots[op["in_slot"]][:, :h * w * c].reshape(n, h, w, c).permute(0, 3, 1, 2)
            y = rb(F.avg_pool2d(x, 2))
            slots[op["out_slot"]] = y.permute(0, 2, 3, 1).reshape(n, -1).contiguous()
        elif k == _lib.PVR_OP_FLATTEN:
            c, h, w, pitch = op["c_in"], op["h_in"], op["w_in"], op["in_pitch"]
            x = slots[op["in_slot"]][:, :h * w * pitch].reshape(n, h * w, pitch)[..., :c]
            emb[:, op["emb_offset"]:op["emb_offset"] + c * h * w] = x.permute(0, 2, 1).reshape(n, -1)
        elif k == _lib.PVR_OP_AVGPOOL:
            c, h, w = op["c_in"], op["h_in"], op["w_in"]
            x = slots[op["in_slot"]][:, :h * w * c].reshape(n, h * w, c)
            emb[:, op["emb_offset"]:op["emb_offset"] + c] = x.mean(1)
        elif k == _lib.PVR_OP_HEAD:
            c, h, w, pitch = op["c_out"], op["h_in"], op["w_in"], op["in_pitch"]
            t = slots[op["in_slot"]][:, :h * w * pitch].reshape(n, h, w, pitch)
            aux = op["_aux"]
            if op.get("act", 0) == 1:  # per-tap partial sums: t[p] = bn(sum_tap Z[p + tap - 1][tap])
                cp = (c + 3) // 4 * 4
                base = 9 * c * cp
                zt = F.pad(t[..., :9 * 2 * c].reshape(n, h, w, 9, 2 * c), (0, 0, 0, 0, 1, 1, 1, 1))
                acc = sum(zt[:, r:r + h, s:s + w, r * 3 + s] for r in range(3) for s in range(3))
                sc1 = aux[base + 2 * c:base + 4 * c]
                bi1 = aux[base + 4 * c:base + 6 * c]
                t = acc * sc1 + bi1
                t = torch.cat([t[..., :c].relu(), t[..., c:]], -1)
                w2 = aux[:base].reshape(3, 3, c, cp)[..., :c].permute(3, 2, 0, 1)  # (r, s, ci, co) -> (co, ci, r, s)
            else:
                base = c * 9 * c
                w2 = aux[:base].reshape(c, 3, 3, c).permute(0, 3, 1, 2)
            a = t[..., :c].permute(0, 3, 1, 2)
            idn = t[..., c:2 * c].permute(0, 3, 1, 2)
            s2, b2 = aux[base:base + c], aux[base + c:base + 2 * c]
            y = F.conv2d(a, w2, padding=1) * s2[None, :, None, None] + b2[None, :, None, None] + idn
            emb[:, op["emb_offset"]:op["emb_offset"] + c * h * w] = y.relu().reshape(n, -1)
    return (emb, slots) if return_slots else emb
