"""BC steps/s under torchrun (data parallel, NCCL): the eager step against a whole-step CUDA graph with the collectives
captured. EXPERIMENT, not on by default: BCTrainer only replays the step from a graph on one GPU. The one 2-GPU trial of
round 1 (use_graph forced on with world 2, NCCL all-reduces inside torch.cuda.graph) did not finish within 200 s and
was killed by its timeout; the cause (NCCL capture vs. the side-stream fork/join of the LSTM wavefront vs. the
watchdog thread) is not isolated yet. BCTrainer ignores use_graph=True for world > 1 until that is understood.
What the trial had changed in BCTrainer._graph_step (reverted; re-apply for the experiment): capture with
`torch.cuda.graph(g, capture_error_mode="thread_local")`, all-reduce `self._gloss` inside the capture when world > 1,
all-reduce the loss of the three eager warm-up steps too, and allow `use_graph` for world > 1 when
`batch_size % world == 0`. Things to try first: TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group (PyTorch's
recipe for capturing NCCL collectives), PVR_LSTM_CHUNKS=1 inside the capture (no side-stream fork/join), and a
side-stream warm-up before the capture.
Usage: torchrun --nproc-per-node N tools/profile_bc_ddp.py [STEPS] [eager|graph ...]   (default: eager only)"""
import os
import random
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvr_habitat_b200.bc import BCTrainer  # noqa: E402
from pvr_habitat_b200.models import PolicyNet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
obs, action, done, _ = bench.bc_dataset()
for mode in (sys.argv[2:] or ["eager"]):
    torch.manual_seed(1)
    random.seed(1)
    net = PolicyNet((2048,), 3, batch_norm=True).cuda().train()
    tr = BCTrainer(net, obs, action, done, 128, 64, 10 ** 9, process_group=dist.group.WORLD, use_graph=(mode == "graph"))
    if mode == "graph":
        tr.use_graph = True  # force the experiment: BCTrainer itself keeps world > 1 on the eager step
    for _ in range(8):  # eager warm-up, graph capture, first replays
        tr.step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    if dist.get_rank() == 0:
        print(f"world {dist.get_world_size()} {mode} (graph in use: {tr._graph is not None}): "
              f"{1e3 * (t1 - t0) / steps:.2f} ms/step = {steps / (t1 - t0):.1f} steps/s, loss {float(tr.last_loss):.6f}",
              flush=True)
dist.destroy_process_group()
