"""Data-parallel plumbing (one process per GPU, torch.distributed / NCCL). The reference has no distributed code
(SURVEY.md D7); the parity target is the single-process reference at the GLOBAL batch:

* embedding pass: observations are cut into contiguous blocks per rank, no collective;
* BC training: every rank draws the SAME seeded `sample_with_minimum_distance` and keeps the sequences
  `starting_i[rank*B/G:(rank+1)*B/G]`; the loss is scaled by 1/(T*B_global) and gradients are SUM all-reduced, so the
  update equals the reference's mean over the global batch; BatchNorm1d statistics are computed from all-reduced
  per-feature sums (synchronised batch statistics, global count).
"""
import os

import torch


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (blocks differ by at most one item)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_starts(starting_i, rank, world):
    """Sequences of the global BC batch trained by `rank` (B must be divisible by the world size so that every rank
    runs the same shapes)."""
    b = len(starting_i)
    if b % world:
        raise ValueError(f"batch_size {b} is not divisible by the world size {world}")
    per = b // world
    return list(starting_i[rank * per:(rank + 1) * per])


def resolve_group(process_group=None):
    """The process group the data-parallel collectives run on, or None for a single process. `None` under an
    initialised torch.distributed with more than one rank — the normal torchrun idiom — means the default (WORLD)
    group, never "no collectives"."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return None
    group = process_group if process_group is not None else torch.distributed.group.WORLD
    return group if torch.distributed.get_world_size(group) > 1 else None


def attach(policy, process_group, global_rows, broadcast=True):
    """Make `policy` (pvr_habitat_b200.models.PolicyNet) synchronise BatchNorm sums and gradients over the group.
    Parameters and buffers are broadcast from the group's first rank, so the replicas start identical whatever the
    seeds of the ranks were."""
    group = resolve_group(process_group)
    policy.process_group = group
    policy.global_rows = global_rows if group is not None else None
    if group is not None and broadcast:
        src = torch.distributed.get_global_rank(group, 0)
        with torch.no_grad():
            for t in list(policy.parameters()) + list(policy.buffers()):
                torch.distributed.broadcast(t.data, src=src, group=group)
    return policy


def allreduce_sum_(t, group=None):
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, group=group)
    return t
