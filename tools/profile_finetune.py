"""Tiny driver for ncu / timing: a few finetune steps at the bench shape (T = 100, B = 16, 64x64x6).
Usage: profile_finetune.py STEPS"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvr_habitat_b200.bc import BCTrainer  # noqa: E402
from pvr_habitat_b200.models import PolicyNetWithConv  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
obs, action, done = bench.finetune_dataset()
torch.manual_seed(1)
random.seed(1)
net = PolicyNetWithConv((64, 64, 6), 3, batch_norm=True).cuda().train()
tr = BCTrainer(net, obs, action, done, 16, 100, 10 ** 9)
for _ in range(6):
    tr.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    tr.step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{steps / dt:.1f} steps/s, {1e3 * dt / steps:.2f} ms/step, loss {float(tr.last_loss):.4f}")
