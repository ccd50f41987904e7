"""Batch-1 / batch-2 encoder latency (the online rollout path): host wall time per `net.embed` call
including the synchronisation a rollout step needs. Usage: latency.py NAME [N_OBS]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "moco_aug"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = bench.build_net(name, torch.device("cuda", 0))
obs = torch.from_numpy(bench.make_observations(n, 1, 3)).cuda()
out = torch.empty(n, net.out_size, device="cuda")
for _ in range(20):
    net.embed(obs, 1, out)
torch.cuda.synchronize()
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    net.embed(obs, 1, out)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
ts.sort()
print(f"{name} batch {n}: median {1e3 * ts[len(ts) // 2]:.3f} ms, p10 {1e3 * ts[20]:.3f} ms, p90 {1e3 * ts[180]:.3f} ms")
