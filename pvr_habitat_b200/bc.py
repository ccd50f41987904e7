"""The BC training hot loop of main_bc_2.py:186-227 (identical in main_bc_1.py:193-234) on the CUDA policy path.

Per step: `sample_with_minimum_distance` (python `random`, same draws as the reference) -> windows of T consecutive
samples wrapping modulo n -> PolicyNet forward -> mean cross-entropy -> `scheduler.step()` BEFORE the update (the
reference's order, so lr_k = lr0 (1 - k / max_epochs)) -> backward -> global-norm clip + RMSprop in one fused kernel.
The dataset lives in HBM: the (T, B) index matrix is the only per-step host->device traffic; the
reference's host gather + H2D of T*B*D floats is kept as `host_batches=True` for the end-to-end measurement.
Environment rollouts / evaluation (simulator) are out of scope; the stats schema (`frames`, `training_loss`,
`gradient_norm`) is the reference's (main_bc_2.py:165-180, 244-246).
"""
import os

import numpy as np
import torch

from . import parallel
from .models import PolicyNet, PolicyNetWithConv, bc_loss
from .optim import FusedAdam, FusedRMSprop
from .utils_bc import sample_with_minimum_distance, window_indices


class BCTrainer:
    def __init__(self, actor_model, obs, action, done, batch_size, unroll_length, max_frames, learning_rate=1e-4,
                 alpha=0.99, epsilon=1e-5, momentum=0, max_grad_norm=40.0, optimizer="rmsprop", process_group=None,
                 host_batches=False, use_graph=None):
        assert isinstance(actor_model, PolicyNet)
        self.model = actor_model
        self.device = actor_model.device
        self.T, self.B = unroll_length, batch_size
        self.rank, self.world = 0, 1
        self.group = parallel.resolve_group(process_group)  # None only when this process trains alone
        if self.group is not None:
            self.rank = torch.distributed.get_rank(self.group)
            self.world = torch.distributed.get_world_size(self.group)
        self.n_samples = len(action)
        self.host_batches = host_batches
        self._next = None
        if host_batches:
            # the reference's data path (dataset in host memory, main_bc_2.py:196-203), pipelined: the rows of step
            # i+1 are gathered into pinned staging buffers and copied on a side stream while step i runs on the GPU
            self.obs = torch.from_numpy(np.ascontiguousarray(np.asarray(obs)))
            self.action = torch.from_numpy(np.ascontiguousarray(np.asarray(action))).long()
            self.done = torch.from_numpy(np.ascontiguousarray(np.asarray(done)))
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = None
            self._stage_i = 0
        else:
            # (a tensor that is already on the device — the table main_bc_1 fills straight from the encoder — is kept)
            to_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).to(self.device)  # noqa: E731
            self.obs, self.action, self.done = to_dev(obs), to_dev(action).long(), to_dev(done)
        self.global_rows = unroll_length * batch_size
        if self.world > 1:
            parallel.attach(actor_model, self.group, self.global_rows)
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
        # Whole-step CUDA graph (single process, RMSprop): forward, loss, backward, clip + update are ~100 launches
        # issued through ctypes in ~4.8 ms of host time against ~4 ms of GPU time; replayed from one graph the step is
        # GPU bound. The batch is copied into static buffers, the learning rate lives in device memory.
        # the whole step replays from one CUDA graph when everything in it is a stream-ordered launch: RMSprop with the
        # learning rate in device memory, and — data parallel — collectives issued through our own NCCL communicator
        # (parallel.Comm.capturable; torch.distributed collectives of other backends are not captured)
        # (PolicyNetWithConv included: its per-step torch temporaries come from the graph's private pool;
        # PVR_FINETUNE_GRAPH=0 is the A/B switch: 6.8 ms eager against 5.8 ms replayed per step at T = 100, B = 16)
        capturable = optimizer == "rmsprop" and \
            (not isinstance(actor_model, PolicyNetWithConv) or os.environ.get("PVR_FINETUNE_GRAPH") != "0") and \
            (self.world == 1 or (actor_model.comm is not None and actor_model.comm.capturable))
        if use_graph is None:
            use_graph = capturable
        self.use_graph = bool(use_graph) and capturable
        self._graph = None
        self._eager_steps = 0

    def make_batch(self, starting_i):
        idx = window_indices(starting_i, self.T, self.n_samples)  # (T, B_local)
        if self.host_batches:
            flat = torch.from_numpy(np.ascontiguousarray(idx.reshape(-1)))
            shape = tuple(idx.shape)
            if self._stage is None or self._stage[0][0].shape[0] != flat.numel():
                self._stage = [tuple(torch.empty((flat.numel(),) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
                                     for t in (self.obs, self.action, self.done)) for _ in range(2)]
                self._stage_ev = [torch.cuda.Event() for _ in range(2)]
            k = self._stage_i
            self._stage_i ^= 1
            self._stage_ev[k].synchronize()  # the copy that last used this staging buffer has completed
            out = []
            with torch.cuda.stream(self._copy_stream):
                for src, pin in zip((self.obs, self.action, self.done), self._stage[k]):
                    torch.index_select(src, 0, flat, out=pin)  # multi-threaded host gather into pinned memory
                    out.append(pin.to(self.device, non_blocking=True).reshape(shape + tuple(src.shape[1:])))
                self._stage_ev[k].record(self._copy_stream)
            return tuple(out) + (self._stage_ev[k],)
        ix = torch.from_numpy(idx).to(self.device, non_blocking=True)
        return self.obs[ix], self.action[ix], self.done[ix], None

    def _draw(self):
        starting_i = sample_with_minimum_distance(n=self.n_samples, k=self.B, d=self.T)
        mine = parallel.shard_starts(starting_i, self.rank, self.world)
        return self.make_batch(mine) + (len(mine),)

    def step(self):
        """One optimisation step; returns the (device) loss of the GLOBAL batch."""
        o, a, d, ready, n_mine = self._next if self._next is not None else self._draw()
        self._next = None
        if ready is not None:
            torch.cuda.current_stream(self.device).wait_event(ready)
            for t in (o, a, d):
                t.record_stream(torch.cuda.current_stream(self.device))
        if self.use_graph:
            loss = self._graph_step(o, a, d, n_mine)
        else:
            loss = self._step_body(o, a, d, n_mine, None)
            if self.world > 1:
                loss = self.model.comm.all_reduce(loss.detach().clone())
        self.frames += self.T * self.B
        self.last_loss = loss.detach()
        if self.host_batches:  # gather + copy the next batch while the GPU works on this one (same draw order)
            self._next = self._draw()
        return self.last_loss

    def _step_body(self, o, a, d, n_mine, lr_tensor, state=None):
        if state is None:
            state = tuple(s.to(self.device) for s in self.model.initial_state(batch_size=n_mine))
        # (the action sampled in training mode is unused here: our policies can skip the draw)
        kw = {"sample_action": False} if getattr(self.model, "accepts_sample_action", False) else {}
        output, _ = self.model(dict(obs=o, done=d), state, **kw)
        loss = bc_loss(output['policy_logits'], a, global_rows=self.global_rows)
        if lr_tensor is None:
            self.scheduler.step()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step(lr_tensor=lr_tensor)
        return loss

    def _graph_step(self, o, a, d, n_mine):
        if self._graph is None and self._eager_steps < 3:  # warm-up: lazy allocations, kernel attributes, LSTM graphs
            self._eager_steps += 1
            loss = self._step_body(o, a, d, n_mine, None)
            return self.model.comm.all_reduce(loss.detach().clone()) if self.world > 1 else loss
        self.scheduler.step()  # before the update, like the reference: lr_k = lr0 (1 - k / max_epochs)
        lr = float(self.optimizer.param_groups[0]["lr"])
        if self._graph is None:
            self._go, self._ga, self._gd = o.clone(), a.clone(), d.clone()
            # the learning rate of THIS step travels as a kernel argument of the fill (stream ordered): no pinned
            # scalar that a host running ahead of the replay could overwrite
            self._lr_dev = torch.full((1,), lr, dtype=torch.float32, device=self.device)
            self._gstate = tuple(s.to(self.device) for s in self.model.initial_state(batch_size=n_mine))
            self._gws = self.model._workspace(self.T, n_mine)  # the graph holds raw pointers into this workspace
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):  # records the step, does not run it: the replay below performs it
                loss = self._step_body(self._go, self._ga, self._gd, n_mine, self._lr_dev, self._gstate)
                self._gloss = loss.detach().clone()
                if self.world > 1:
                    self.model.comm.all_reduce(self._gloss)
            self._graph = g
        else:
            self._go.copy_(o, non_blocking=True)
            self._ga.copy_(a, non_blocking=True)
            self._gd.copy_(d, non_blocking=True)
            self._lr_dev.fill_(lr)
            self.optimizer.count_replayed_step()  # `state[p]['step']` of a checkpoint stays the true step count
        self._graph.replay()
        return self._gloss.clone()

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


# ---------------------------------------------------------------------------------------------- checkpoint schema
CHECKPOINT_KEYS = ("embedding_model_state_dict", "actor_model_state_dict", "actor_model_optimizer_state_dict",
                   "scheduler_state_dict", "flags")


def save_checkpoint(save_path, embedding_model, actor_model, optimizer, scheduler, flags, stats=None):
    """The `.tar` (+ `.pickle` statistics) pair of main_bc_2.py:250-258 / main_bc_1.py:259-267, same keys and the same
    state_dict layouts (PolicyNet, FusedRMSprop and LambdaLR mirror the reference's modules), so runs interchange with
    the reference in both directions. `flags` is an argparse namespace or a dict."""
    import pickle
    if stats is not None:
        with open(save_path + '.pickle', 'wb') as fh:
            pickle.dump(stats, fh, protocol=pickle.HIGHEST_PROTOCOL)
    ck = {
        'actor_model_state_dict': actor_model.state_dict(),
        'actor_model_optimizer_state_dict': optimizer.state_dict(),
        'scheduler_state_dict': scheduler.state_dict(),
        'flags': dict(flags) if isinstance(flags, dict) else vars(flags),
    }
    if embedding_model is not None:  # main_bc_finetune.py:232-238 has no separate encoder
        ck = {'embedding_model_state_dict': embedding_model.state_dict(), **ck}
    torch.save(ck, save_path + '.tar')


def load_checkpoint(path, embedding_model=None, actor_model=None, optimizer=None, scheduler=None):
    """Restore whichever objects are given from a `.tar` written by `save_checkpoint` or by the reference."""
    ck = torch.load(path, map_location='cpu', weights_only=False)
    missing = [k for k in CHECKPOINT_KEYS if k not in ck and not (k == "embedding_model_state_dict" and
                                                                    embedding_model is None)]
    if missing:
        raise KeyError(f"{path}: not a BC checkpoint (missing {missing})")
    if embedding_model is not None:
        embedding_model.load_state_dict(ck['embedding_model_state_dict'])
    if actor_model is not None:
        actor_model.load_state_dict(ck['actor_model_state_dict'])
    if optimizer is not None:
        optimizer.load_state_dict(ck['actor_model_optimizer_state_dict'])
    if scheduler is not None:
        scheduler.load_state_dict(ck['scheduler_state_dict'])
    return ck['flags']
