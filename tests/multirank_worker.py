"""Worker of tests/test_gpu_multirank.py, launched once per rank by torch.distributed.run.

    python -m torch.distributed.run --nproc-per-node 2 ... tests/multirank_worker.py <case> <out.json>

Trains the SAME seeded problem data-parallel over the ranks through the product API (BCTrainer built under an
initialised torch.distributed WITHOUT an explicit process_group — the normal torchrun idiom) and writes the loss /
gradient-norm trace of the global batch plus parameter checksums. The launching test compares it with the
single-process run of the same global batch (the reference is single process: SURVEY.md D7).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem(case):
    """(model, dataset, T, B, steps) of a case; identical in every process that calls it (seeded)."""
    from oracle import restate_policy as rp
    from pvr_habitat_b200.models import PolicyNet, PolicyNetWithConv
    if case == "policy":
        T, B, D, n, steps = 16, 8, 256, 2048, 10
        obs, action, done, _ = rp.synthetic_bc_data(n, D, 3, 5)
        torch.manual_seed(11)
        net = PolicyNet((D,), 3, batch_norm=True)
    elif case == "finetune":
        T, B, steps = 6, 4, 6
        data, action = rp.synthetic_frame_trajectories(6, 20, 9)
        obs = np.concatenate(data["obs"])
        done = np.concatenate(data["done"])
        torch.manual_seed(12)
        net = PolicyNetWithConv((64, 64, 6), 3, batch_norm=True)
    else:
        raise ValueError(case)
    return net, obs, action, done, T, B, steps


def train(case, device, keep_state=False, **kw):
    import random
    from pvr_habitat_b200.bc import BCTrainer
    net, obs, action, done, T, B, steps = problem(case)
    net = net.to(device).train()
    random.seed(3)
    tr = BCTrainer(net, obs, action, done, B, T, max_frames=steps * T * B * 2, **kw)
    losses, norms = [], []
    for _ in range(steps):
        losses.append(float(tr.step().item()))
        norms.append(float(tr.gradient_norm().item()))
    sums = {k: float(v.double().sum()) for k, v in net.state_dict().items() if v.is_floating_point()}
    res = dict(loss=losses, grad_norm=norms, param_sums=sums, world=tr.world)
    if keep_state:
        res["state"] = {k: v.detach().float().cpu() for k, v in net.state_dict().items() if v.is_floating_point()}
    return res


def main():
    case, out = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_gpu = torch.cuda.device_count()
    if n_gpu >= world:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:  # one GPU: both ranks share it, collectives through gloo (exercises the same product code path)
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    if rank != 0:  # a different seed on the other ranks: the replicas must be made identical by the broadcast
        torch.manual_seed(1234 + rank)
    res = train(case, torch.device("cuda", torch.cuda.current_device()), keep_state=True)
    assert res["world"] == world
    state = res.pop("state")
    gathered = [None] * world
    dist.all_gather_object(gathered, res["param_sums"])
    if rank == 0:
        res["backend"] = dist.get_backend()
        res["replica_param_sums"] = gathered
        with open(out, "w") as f:
            json.dump(res, f)
        torch.save(state, out + ".state.pt")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
