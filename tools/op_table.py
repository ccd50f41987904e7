"""Per-op table of one embedding pass, timed live with CUDA events (pvr_encoder_forward_timed), next to the
algorithmic FLOPs and HBM bytes of every op. Usage: op_table.py NAME N_FRAMES [REPS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "moco_aug"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
net = bench.build_net(name, torch.device("cuda", 0))
net.max_images_per_pass = n  # one pass: the encoder stays bound to all n frames for the timed forwards below
obs = torch.from_numpy(bench.make_observations(n, 1, 3)).cuda()
out = net.embed(obs, 1)
enc = net.encoder()
acc = None
for _ in range(reps):
    net.transforms.run(obs, 1, enc.slot0, enc.input_format, True)
    ms = enc.forward_timed(out, net.out_size)
    acc = ms if acc is None else [a + b for a, b in zip(acc, ms)]
ms = [a / reps for a in acc]
KIND = {1: "conv", 2: "maxpool", 3: "avgpool", 4: "head", 5: "flatten"}
print(f"{name}, {n} frames, mean of {reps} passes")
print(f"{'op':>3} {'kind':8} {'shape':28} {'us':>8} {'TF/s':>7} {'HBM MB':>8} {'GB/s':>7}")
tot = 0.0
for i, (m, t) in enumerate(zip(enc.op_meta, ms)):
    k = KIND.get(m["kind"], "?")
    flops = m.get("flops_per_image", 0) * n
    b_in = m["h_in"] * m["w_in"] * m["c_in"] * 2 * n
    b_out = m.get("h_out", 0) * m.get("w_out", 0) * m.get("c_out", 0) * 2 * n if m.get("out_slot", -1) >= 0 else 0
    b_res = b_out if m.get("res_slot", -1) >= 0 else 0
    b_w = m.get("k_pad", 0) * m.get("n_pad", 0) * 2
    mb = (b_in + b_out + b_res + b_w) / 1e6
    shape = f"{m['c_in']}->{m.get('c_out', 0)} {m.get('r', 1)}x{m.get('s', 1)}/{m.get('stride_h', 1)} @{m.get('h_out', 0)}" + \
        ("+res" if b_res else "")
    print(f"{i:3d} {k:8} {shape:28} {t * 1e3:8.1f} {flops / t / 1e9 if t > 0 else 0:7.0f} {mb:8.1f} {mb / t if t > 0 else 0:7.0f}")
    tot += t
print(f"total {tot * 1e3:.1f} us")
# ---- by category
cat = {}
for m, t in zip(enc.op_meta, ms):
    if m["kind"] != 1:
        k = KIND.get(m["kind"], "?")
    elif m.get("r", 1) == 7:
        k = "stem (+pool)"
    elif m.get("r", 1) == 3:
        k = f"3x3 @{m['h_out']}"
    elif m.get("res_slot", -1) >= 0:
        k = f"1x1+res @{m['h_out']}"
    else:
        k = f"1x1 @{m['h_out']}"
    cat[k] = cat.get(k, 0.0) + t
for k, v in sorted(cat.items(), key=lambda kv: -kv[1]):
    print(f"  {k:16s} {v * 1e3:8.1f} us {100 * v / tot:5.1f} %")
