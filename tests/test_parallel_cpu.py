"""N > 1 host logic on CPU (gloo, world_size 2): sharding of observations / BC sequences, and the data-parallel
identity the BC path relies on — local losses scaled by 1/(T*B_global), SUM all-reduce of gradients and of the
BatchNorm (sum, sum of squares) — reproduces the single-process global-batch gradient. The arithmetic here is the
oracle's (test infrastructure); the CUDA path is exercised under `gpurun --gpus N`."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate_policy as rp
from pvr_habitat_b200 import parallel, utils_bc


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_starts_partitions_the_global_batch():
    random.seed(0)
    starts = utils_bc.sample_with_minimum_distance(1000, 16, 20)
    parts = [parallel.shard_starts(starts, r, 4) for r in range(4)]
    assert sum(parts, []) == starts and all(len(p) == 4 for p in parts)
    with pytest.raises(ValueError):
        parallel.shard_starts(starts, 0, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    T, B, D = 4, 4, 32
    obs, action, done, _ = rp.synthetic_bc_data(256, D, 3, 1)
    sd = rp.init_policy_state(D, 3, False, 3)
    random.seed(5)  # every rank draws the same global sample
    starts = utils_bc.sample_with_minimum_distance(256, B, T)
    mine = parallel.shard_starts(starts, rank, world)
    o, a, d = rp.make_batch(obs, action, done, mine, T)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("baseline.")}
    zero = (torch.zeros(2, len(mine), 1024), torch.zeros(2, len(mine), 1024))
    logits, _, _ = rp.policy_forward({**sd, **params}, torch.from_numpy(o), torch.from_numpy(d), zero, False)
    nll = torch.nn.functional.nll_loss(torch.log_softmax(logits.flatten(0, 1), -1),
                                       torch.from_numpy(a).flatten().long(), reduction="sum")
    loss = nll / (T * B)  # scaled by the GLOBAL row count
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in params.values()])
    parallel.allreduce_sum_(flat)
    total = loss.detach().clone()
    parallel.allreduce_sum_(total)
    # BatchNorm statistics from all-reduced sums
    x = torch.from_numpy(o).flatten(0, 1).double()
    sums = torch.stack([x.sum(0), (x * x).sum(0)])
    parallel.allreduce_sum_(sums)
    if rank == 0:
        torch.save(dict(flat=flat, loss=total, sums=sums), out)
    dist.destroy_process_group()


def test_two_rank_data_parallel_equals_global_batch(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    T, B, D = 4, 4, 32
    obs, action, done, _ = rp.synthetic_bc_data(256, D, 3, 1)
    sd = rp.init_policy_state(D, 3, False, 3)
    random.seed(5)
    starts = utils_bc.sample_with_minimum_distance(256, B, T)
    o, a, d = rp.make_batch(obs, action, done, starts, T)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("baseline.")}
    zero = (torch.zeros(2, B, 1024), torch.zeros(2, B, 1024))
    logits, _, _ = rp.policy_forward({**sd, **params}, torch.from_numpy(o), torch.from_numpy(d), zero, False)
    loss = rp.bc_loss(logits, torch.from_numpy(a))
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in params.values()])
    assert abs(float(got["loss"]) - float(loss)) < 1e-6
    torch.testing.assert_close(got["flat"], flat, atol=1e-7, rtol=1e-4)
    x = torch.from_numpy(o).flatten(0, 1).double()
    torch.testing.assert_close(got["sums"], torch.stack([x.sum(0), (x * x).sum(0)]))


# ------------------------------------------------------------------------------------------------ product plumbing
def _attach_worker(rank, world, port, out):
    """The product's attach / Comm path on gloo (CPU tensors): the normal torchrun idiom — torch.distributed is
    initialised and NO process group is passed."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvr_habitat_b200.models import PolicyNet
    torch.manual_seed(100 + rank)  # different seeds: the replicas must be made identical by the broadcast
    net = PolicyNet((16,), 3, batch_norm=True)
    before = float(net.policy.weight.double().sum())
    group = parallel.resolve_group(None)
    assert group is not None and dist.get_world_size(group) == world
    parallel.attach(net, None, global_rows=48)
    assert net.process_group is not None and net.comm is not None and net.global_rows == 48
    assert net.comm.world == world and not net.comm.capturable  # gloo: torch.distributed fallback, never captured
    t = torch.full((5,), float(rank + 1), dtype=torch.float64)
    net.comm.all_reduce(t)
    sums = {k: float(v.double().sum()) for k, v in net.state_dict().items() if v.is_floating_point()}
    gathered = [None] * world
    dist.all_gather_object(gathered, (before, sums, t.tolist()))
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_attach_resolves_the_default_group_and_broadcasts(tmp_path):
    """Round-1 advisor finding: with process_group=None under an initialised torch.distributed the collectives were
    skipped and every rank trained alone. `attach` must pick WORLD, create the Comm and make the replicas identical."""
    out = str(tmp_path / "attach.pt")
    mp.spawn(_attach_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (b0, s0, t0), (b1, s1, t1) = torch.load(out)
    assert b0 != b1                      # the ranks started from different initialisations ...
    assert s0 == s1                      # ... and hold rank 0's parameters / buffers after attach
    assert abs(s0["policy.weight"] - b0) < 1e-12
    assert t0 == t1 == [3.0] * 5         # all_reduce is a SUM over the ranks


def test_resolve_group_single_process():
    assert parallel.resolve_group(None) is None  # no torch.distributed: one process, no collectives
