"""Tiny driver for ncu: one (or a few) embedding passes of a workload. Usage: profile_embed.py NAME N_OBS N_FRAMES REPS"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "moco_aug"
n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
net = bench.build_net(name, torch.device("cuda", 0))
obs = torch.from_numpy(bench.make_observations(n_obs, n_frames, 3)).cuda()
for _ in range(reps):
    out = net.embed(obs, n_frames)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()))
