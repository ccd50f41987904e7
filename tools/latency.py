"""Online rollout latency (src/test_model.py:11-17 + EmbeddingWrapper.observation, src/embeddings.py:441-444):
host wall time of one step = batch-N `net.embed` of a uint8 observation -> one PolicyNet step (T = 1) -> `action` on the
host, with the synchronisation a rollout needs. Both halves replay from CUDA graphs after their warm-up calls.
Usage: latency.py NAME [N_OBS]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvr_habitat_b200.models import PolicyNet  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "moco_aug"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = bench.build_net(name, torch.device("cuda", 0))
obs = torch.from_numpy(bench.make_observations(n, 1, 3)).cuda()
out = torch.empty(n, net.out_size, device="cuda")


def median(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2], 1e3 * ts[reps // 10], 1e3 * ts[reps * 9 // 10]


m = median(lambda: net.embed(obs, 1, out))
print(f"{name} batch {n}: embed median {m[0]:.3f} ms, p10 {m[1]:.3f} ms, p90 {m[2]:.3f} ms")

torch.manual_seed(0)
policy = PolicyNet((net.out_size,), 3, batch_norm=True).cuda().eval()
done = torch.zeros(1, n, dtype=torch.bool)
state = [tuple(s.cuda() for s in policy.initial_state(n))]


def rollout_step():
    net.embed(obs, 1, out)
    with torch.no_grad():
        o, state[0] = policy(dict(obs=out.view(1, n, -1), done=done), state[0])
    return o["action"].cpu()  # the environment consumes the action on the host


for rows, label in ((8, "graph replay"), (0, "eager")):
    policy.rollout_graph_rows = rows
    m = median(rollout_step)
    print(f"{name} batch {n}: embed + policy step + action on host ({label}): median {m[0]:.3f} ms, p10 {m[1]:.3f} ms, "
          f"p90 {m[2]:.3f} ms")
