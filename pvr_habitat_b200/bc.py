"""The BC training hot loop of main_bc_2.py:186-227 (identical in main_bc_1.py:193-234) on the CUDA policy path.

Per step: `sample_with_minimum_distance` (python `random`, same draws as the reference) -> windows of T consecutive
samples wrapping modulo n -> PolicyNet forward -> mean cross-entropy -> `scheduler.step()` BEFORE the update (the
reference's order, so lr_k = lr0 (1 - k / max_epochs)) -> backward -> global-norm clip + RMSprop in one fused kernel.
The dataset lives in HBM: the (T, B) index matrix is the only per-step host->device traffic; the
reference's host gather + H2D of T*B*D floats is kept as `host_batches=True` for the end-to-end measurement.
Environment rollouts / evaluation (simulator) are out of scope; the stats schema (`frames`, `training_loss`,
`gradient_norm`) is the reference's (main_bc_2.py:165-180, 244-246).
"""
import numpy as np
import torch

from . import parallel
from .models import PolicyNet, bc_loss
from .optim import FusedAdam, FusedRMSprop
from .utils_bc import sample_with_minimum_distance, window_indices


class BCTrainer:
    def __init__(self, actor_model, obs, action, done, batch_size, unroll_length, max_frames, learning_rate=1e-4,
                 alpha=0.99, epsilon=1e-5, momentum=0, max_grad_norm=40.0, optimizer="rmsprop", process_group=None,
                 host_batches=False):
        assert isinstance(actor_model, PolicyNet)
        self.model = actor_model
        self.device = actor_model.device
        self.T, self.B = unroll_length, batch_size
        self.rank, self.world = 0, 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.rank = torch.distributed.get_rank(process_group)
            self.world = torch.distributed.get_world_size(process_group)
        self.group = process_group
        self.n_samples = len(action)
        self.host_batches = host_batches
        if host_batches:  # the reference's data path: numpy arrays on the host, gathered and copied every step
            self.obs, self.action, self.done = np.asarray(obs), np.asarray(action), np.asarray(done)
        else:
            self.obs = torch.as_tensor(np.asarray(obs)).to(self.device)
            self.action = torch.as_tensor(np.asarray(action)).long().to(self.device)
            self.done = torch.as_tensor(np.asarray(done)).to(self.device)
        self.global_rows = unroll_length * batch_size
        if self.world > 1:
            parallel.attach(actor_model, process_group, self.global_rows)
        if optimizer == "rmsprop":
            self.optimizer = FusedRMSprop(actor_model.parameters(), lr=learning_rate, momentum=momentum, eps=epsilon,
                                          alpha=alpha, max_grad_norm=max_grad_norm)
        else:
            self.optimizer = FusedAdam(actor_model.parameters(), lr=learning_rate, eps=epsilon,
                                       max_grad_norm=max_grad_norm)
        max_epochs = max_frames // (unroll_length * batch_size) + 1
        self.max_epochs = max_epochs
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lambda epoch: 1 - epoch / max_epochs)
        self.frames = 0
        self.last_loss = None

    def make_batch(self, starting_i):
        idx = window_indices(starting_i, self.T, self.n_samples)  # (T, B_local)
        if self.host_batches:
            o = torch.from_numpy(self.obs[idx]).to(self.device, non_blocking=True)
            a = torch.from_numpy(self.action[idx]).to(self.device, non_blocking=True)
            d = torch.from_numpy(self.done[idx]).to(self.device, non_blocking=True)
            return o, a, d
        ix = torch.from_numpy(idx).to(self.device, non_blocking=True)
        return self.obs[ix], self.action[ix], self.done[ix]

    def step(self):
        """One optimisation step; returns the (device) loss of the GLOBAL batch."""
        starting_i = sample_with_minimum_distance(n=self.n_samples, k=self.B, d=self.T)
        mine = parallel.shard_starts(starting_i, self.rank, self.world)
        o, a, d = self.make_batch(mine)
        state = tuple(s.to(self.device) for s in self.model.initial_state(batch_size=len(mine)))
        output, _ = self.model(dict(obs=o, done=d), state)
        loss = bc_loss(output['policy_logits'], a, global_rows=self.global_rows)
        self.scheduler.step()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        if self.world > 1:
            loss = loss.detach().clone()
            torch.distributed.all_reduce(loss, group=self.group)
        self.frames += self.T * self.B
        self.last_loss = loss.detach()
        return self.last_loss

    def gradient_norm(self):
        return self.optimizer.gradient_norm()


def train_bc(actor_model, obs, action, done, batch_size, unroll_length, max_frames, eval_frequency=200, **kw):
    """Run the loop to `max_frames`; returns the reference's stats dict (without the rollout keys)."""
    tr = BCTrainer(actor_model, obs, action, done, batch_size, unroll_length, max_frames, **kw)
    stats = {"frames": [0], "training_loss": [float("nan")], "gradient_norm": [float("nan")]}
    for frames in range(0, max_frames, batch_size * unroll_length):
        epoch = frames // (batch_size * unroll_length)
        loss = tr.step()
        if (epoch + 1) % eval_frequency == 0:
            stats["frames"].append(frames)
            stats["training_loss"].append(float(loss.item()))
            stats["gradient_norm"].append(float(tr.gradient_norm().item()))
    return stats
