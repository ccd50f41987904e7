"""Fused global-norm clip + RMSprop / Adam (multi-tensor CUDA kernels, no host sync).

Replaces, for the BC loop, the sequence of main_bc_2.py:220-227: the per-parameter `.grad.norm(2).item()` statistic
(18 host syncs), `nn.utils.clip_grad_norm_(params, max_grad_norm)` and `torch.optim.RMSprop.step()`. Works with
`torch.optim.lr_scheduler.LambdaLR` (the learning rate is read from `param_groups` at every step, so calling
`scheduler.step()` before `optimizer.step()` like the reference gives lr_k = lr0 * (1 - k / max_epochs)).
"""
import ctypes

import torch

from . import _lib

_MAX = 32  # tensors per kernel launch (pvr_optim_* limit)


class _FusedBase(torch.optim.Optimizer):
    mode = None

    def __init__(self, params, defaults, max_grad_norm=None, process_group=None):
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.process_group = process_group
        self._sumsq = None
        self._norm = None

    def count_replayed_step(self):
        """A step replayed from a captured CUDA graph does not pass through `step()`: keep the per-parameter step
        counters (checkpoint contents, Adam's bias correction at re-capture) in line with the updates performed."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = int(st["step"]) + 1

    def gradient_norm(self):
        """Pre-clip global gradient norm of the last step (device tensor; `.item()` syncs) — the reference's
        `gradient_norm` statistic (main_bc_2.py:220-224)."""
        return self._norm

    @torch.no_grad()
    def step(self, closure=None, lr_tensor=None):
        """lr_tensor: optional 1-element float32 CUDA tensor holding the learning rate (RMSprop only); the update then
        takes no host scalars that change from step to step and can be captured in a CUDA graph."""
        assert closure is None
        lib = _lib.lib()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if self._sumsq is None:
                self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
                self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
            s1, s2 = [], []
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    self._init_state(st, p)
                st["step"] = int(st["step"]) + 1  # checkpoints written by torch >= 1.12 hold `step` as a tensor
                a, b = self._state_tensors(st)
                s1.append(a)
                s2.append(b)
            step = self.state[ps[0]]["step"]
            with torch.cuda.device(dev):
                stream = _lib.current_stream_ptr()
                n = len(ps)
                assert n <= _MAX, "more than 32 parameter tensors in one group"
                VP, I64 = ctypes.c_void_p * n, ctypes.c_int64 * n
                grads = VP(*[p.grad.data_ptr() for p in ps])
                params = VP(*[p.data_ptr() for p in ps])
                st1 = VP(*[t.data_ptr() for t in s1])
                st2 = VP(*[t.data_ptr() if t is not None else None for t in s2])
                sizes = I64(*[p.numel() for p in ps])
                _lib.check(lib.pvr_optim_sumsq(grads, sizes, n, self._sumsq.data_ptr(), stream), "pvr_optim_sumsq")
                h = self._hyper(group)
                if lr_tensor is not None:
                    _lib.check(lib.pvr_optim_step_dev(self.mode, params, grads, st1, st2, sizes, n,
                                                      self._sumsq.data_ptr(), 1.0, float(self.max_grad_norm or 0.0),
                                                      lr_tensor.data_ptr(), h[0], h[1], h[2], step,
                                                      self._norm.data_ptr(), stream), "pvr_optim_step_dev")
                else:
                    _lib.check(lib.pvr_optim_step(self.mode, params, grads, st1, st2, sizes, n, self._sumsq.data_ptr(),
                                                  1.0, float(self.max_grad_norm or 0.0), float(group["lr"]), h[0], h[1],
                                                  h[2], step, self._norm.data_ptr(), stream), "pvr_optim_step")
        return None


class FusedRMSprop(_FusedBase):
    """torch.optim.RMSprop(lr, alpha, eps, momentum=0, centered=False) semantics (main_bc_2.py:80-85)."""
    mode = 0

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False,
                 max_grad_norm=None, process_group=None):
        if momentum != 0 or weight_decay != 0 or centered:
            raise NotImplementedError("FusedRMSprop: momentum / weight_decay / centered are not on the BC path "
                                      "(main_bc_2.py:80-85, src/arguments.py:61-62)")
        # the param_group keys of torch.optim.RMSprop, so that `actor_model_optimizer_state_dict` of a checkpoint
        # (main_bc_2.py:252-258) loads here and a state_dict written here loads into torch.optim.RMSprop
        super().__init__(params, dict(lr=lr, momentum=momentum, alpha=alpha, eps=eps, centered=centered,
                                      weight_decay=weight_decay), max_grad_norm, process_group)

    def _init_state(self, st, p):
        st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _state_tensors(self, st):
        return st["square_avg"], None

    def _hyper(self, group):
        return float(group["alpha"]), 0.0, float(group["eps"])


class FusedAdam(_FusedBase):
    """torch.optim.Adam(lr, betas, eps) semantics (no weight decay, no amsgrad) — the north star's Adam mode."""
    mode = 1

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None, process_group=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps), max_grad_norm, process_group)

    def _init_state(self, st, p):
        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _state_tensors(self, st):
        return st["exp_avg"], st["exp_avg_sq"]

    def _hyper(self, group):
        return float(group["betas"][0]), float(group["betas"][1]), float(group["eps"])
