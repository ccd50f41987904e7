"""Tiny driver for ncu / timing: a few BC steps at the bench shape. Usage: profile_bc.py STEPS [B [T]]
(B = sequences per step: 128 is the bench shape; 32 / 16 are what a rank holds under 4 / 8-way data parallelism)"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvr_habitat_b200.bc import BCTrainer  # noqa: E402
from pvr_habitat_b200.models import PolicyNet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
obs, action, done, _ = bench.bc_dataset()
torch.manual_seed(1)
random.seed(1)
net = PolicyNet((2048,), 3, batch_norm=True).cuda().train()
tr = BCTrainer(net, obs, action, done, B, T, 10 ** 9)
for _ in range(6):  # eager warm-up, graph capture, first replays
    tr.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    tr.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B} T={T} PVR_LSTM_PERSIST_CHUNKS={os.environ.get('PVR_LSTM_PERSIST_CHUNKS', 'default')}: host issue {1e3 * (t1 - t0) / steps:.2f} ms/step, "
      f"wall {1e3 * (t2 - t0) / steps:.2f} ms/step = {steps / (t2 - t0):.1f} steps/s, loss {float(tr.last_loss):.6f}")
