"""Embedding throughput against the number of frames per pass (L2 residency of the activations).
Usage: sweep_pass_size.py NAME SIZES...   (sizes = observations per pass, 1 frame each)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "moco_aug"
sizes = [int(a) for a in sys.argv[2:]] or [32, 64, 128, 256, 512]
net = bench.build_net(name, torch.device("cuda", 0))
total = max(sizes) * 4
obs = torch.from_numpy(bench.make_observations(total, 1, 3)).cuda()
out = torch.empty(total, net.out_size, device="cuda")
for n in sizes:
    net.max_images_per_pass = n
    for _ in range(2):
        net.embed(obs, 1, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        net.embed(obs, 1, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name} pass={n:5d} frames  {total / ms * 1e3:9.0f} frames/s  ({ms:.2f} ms / {total} frames)", flush=True)
