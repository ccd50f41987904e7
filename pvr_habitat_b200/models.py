"""Drop-in for the reference's src/models.py: PolicyNet (and the fused BC loss).

Same constructor, attributes, state_dict keys (`fc.*`, `core.weight_ih_l0` ..., `policy.*`, `baseline.*`), initial
values (the constructor consumes the torch RNG exactly like the reference: orthogonal init for the Linear layers,
nn.LSTM default init) and forward signature `forward(inputs: {'obs','done'}, core_state) -> (dict, core_state)`
(src/models.py:57-89). The arithmetic runs in libpvr_b200: tcgen05 GEMMs for the Linear layers, the LSTM input
projections and the recurrent matmuls, fused cell / BatchNorm / heads / loss kernels. Parameters stay fp32 (master
copies); GEMM operands are bf16 with fp32 accumulation. There is no CPU path.

Layer-wise execution: the done masks depend only on t, so running all T steps of LSTM layer 0 and then layer 1 equals
the reference's per-timestep two-layer call and lets x_t W_ih^T be one (T*B) x 4096 GEMM per layer.
"""
import ctypes
import os

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from ._lib import pvr_gemm_desc, pvr_lstm_bwd, pvr_lstm_fwd


def init(module, weight_init, bias_init, gain=1):
    weight_init(module.weight.data, gain=gain)
    bias_init(module.bias.data)
    return module


def _r64(x):
    return (x + 63) // 64 * 64


def _stream():
    return _lib.current_stream_ptr()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(a, b, out, m, n, k, bias=None, relu=False, res=None, res_mode=0, out_f32=0, split_k=1, n_pad=None, act=0,
         mn=False):
    """out (m x n) = epilogue(a (m x k) @ b (n x k)^T); a, b bf16 row-major with K contiguous.
    mn=True: a is (k x m), b is (k x n) row-major (MN-major operands): out = a^T b, fp32."""
    d = pvr_gemm_desc()
    d.flags = _lib.PVR_GEMM_MN if mn else 0
    d.a, d.lda = a.data_ptr(), a.stride(0)
    d.b, d.ldb = b.data_ptr(), b.stride(0)
    d.out, d.ldo = out.data_ptr(), out.stride(0)
    d.scale = None
    d.bias = bias.data_ptr() if bias is not None else None
    d.res, d.ldr = (res.data_ptr(), res.stride(0)) if res is not None else (None, 0)
    d.m, d.n, d.n_pad, d.k = m, n, n_pad if n_pad is not None else (b.shape[1] if mn else b.shape[0]), k
    d.relu, d.res_mode, d.out_f32, d.split_k, d.act = int(relu), res_mode, out_f32, split_k, act
    _lib.check(_lib.lib().pvr_gemm(ctypes.byref(d), _stream()), "pvr_gemm")


class _Workspace:
    """Device buffers of one (T, B, D) problem, reused across steps."""

    def __init__(self, T, B, D, H, A, device, batch_norm):
        M, Mp, Dp = T * B, _r64(T * B), _r64(D)
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dtype=f32: torch.zeros(*s, dtype=dtype, device=device)  # noqa: E731
        self.T, self.B, self.D, self.H, self.A, self.M, self.Mp, self.Dp = T, B, D, H, A, M, Mp, Dp
        self.X0 = z(M, Dp, dtype=bf)          # BatchNorm output / cast input (K padded with zeros)
        self.H1, self.H2 = z(M, H, dtype=bf), z(M, H, dtype=bf)
        self.XP = [z(M, 4 * H), z(M, 4 * H)]
        self.HL = [z(M, H, dtype=bf), z(M, H, dtype=bf)]
        self.hm = [z(M, H, dtype=bf), z(M, H, dtype=bf)]
        self.gates = [z(M, 4 * H), z(M, 4 * H)]
        self.c_all = [z(M + B, H), z(M + B, H)]
        self.g_tmp = [z(B, 4 * H), z(B, 4 * H)]  # per layer: the two layers run concurrently (wavefront)
        # arrival counters of the persistent recurrence kernels, one buffer per layer of THIS workspace (never shared
        # with another sequence that may run on another stream)
        self.lstm_counters = [torch.zeros(_lib.lstm_counter_bytes(T, B) // 4, dtype=torch.int32, device=device)
                              for _ in range(2)]
        self.h_last = [z(B, H), z(B, H)]
        self.h0 = [z(B, H), z(B, H)]          # persistent copies: stable pointers keep the CUDA-graph cache warm
        self.nd = z(T, B)
        self.logits, self.baseline = z(M, A), z(M)
        self.mean, self.rstd = z(D), z(D)
        self.sums = torch.zeros(2 * D, dtype=torch.float64, device=device)
        # backward
        self.dHL = [z(M, H), z(M, H)]         # fp32 gradients w.r.t. the LSTM layer outputs
        self.dG = [z(M, 4 * H, dtype=bf), z(M, 4 * H, dtype=bf)]
        self.dh_rec, self.dc_rec = [z(B, H), z(B, H)], [z(B, H), z(B, H)]
        self.dZ2, self.dZ1 = z(M, H, dtype=bf), z(M, H, dtype=bf)
        self.dX0 = z(M, Dp, dtype=bf) if batch_norm else None
        # (the weight-gradient GEMMs read dY / X as MN-major operands: no transposed copies, see gemm(mn=True))
        self.dW1p = z(H, Dp) if Dp != D else None


class _PolicyFn(torch.autograd.Function):
    """forward + backward of the whole PolicyNet as one autograd node (the CUDA kernels do not build a graph)."""

    @staticmethod
    def forward(ctx, net, x, notdone, h0, c0, *params):
        ctx.set_materialize_grads(False)
        out = net._forward_cuda(x, notdone, h0, c0)
        ctx.net = net
        ctx.generation = net._generation
        ctx.n_params = len(params)
        ctx.mark_non_differentiable(out[2], out[3])
        return out

    @staticmethod
    def backward(ctx, dlogits, dbaseline, dh, dc):
        if dbaseline is not None:
            raise NotImplementedError("PolicyNet backward: a loss on `baseline` is not part of the BC path "
                                      "(main_bc_2.py:211-214 uses policy_logits only)")
        if dlogits is None:
            return (None,) * (5 + ctx.n_params)
        ctx.net._check_generation(ctx.generation)
        grads = ctx.net._backward_cuda(dlogits.contiguous())
        return (None, None, None, None, None) + tuple(grads)


class PolicyNet(nn.Module):
    def __init__(self, observation_shape, num_actions, batch_norm=False):
        super(PolicyNet, self).__init__()
        self._build_trunk(observation_shape[0], num_actions, batch_norm)

    def _build_trunk(self, obs_size, num_actions, batch_norm):
        observation_shape = (obs_size,)
        init_ = lambda m: init(m, nn.init.orthogonal_,  # noqa: E731
                               lambda x: nn.init.constant_(x, 0), nn.init.calculate_gain('relu'))
        # identical module structure / construction order to src/models.py:22-44 (same RNG stream, same keys)
        self.fc = nn.Sequential(
            init_(nn.Linear(observation_shape[0], 1024)),
            nn.ReLU(),
            init_(nn.Linear(1024, 1024)),
            nn.ReLU(),
        )
        if batch_norm:
            self.fc = nn.Sequential(nn.BatchNorm1d(observation_shape[0]), *list(self.fc))
        self.core = nn.LSTM(1024, 1024, 2)
        init_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0))  # noqa: E731
        self.policy = init_(nn.Linear(1024, num_actions))
        self.baseline = init_(nn.Linear(1024, 1))

        self.batch_norm = bool(batch_norm)
        self.num_actions = num_actions
        self.obs_size = observation_shape[0]
        self._ws = {}
        self._side = None             # second stream of the LSTM wavefront
        self._rollout = {}            # (T, B, device) -> captured evaluation step (see _rollout_step)
        self.rollout_graph_rows = 8   # T*B up to which no-grad eval forwards replay from a CUDA graph (0 = never)
        self._needs_input_grad = False  # PolicyNetWithConv: the input rows are conv features
        self._wb = None          # bf16 weight copies
        self._saved = None
        self._generation = 0     # bumped by every forward: the activations saved for the backward live in ONE set of
        #                          workspace buffers per (T, B), so a backward must follow ITS forward directly
        # data-parallel hooks (set by pvr_habitat_b200.parallel): all-reduce of the BatchNorm sums
        self.process_group = None
        self.comm = None         # parallel.Comm of the group (NCCL through the C ABI, or torch.distributed)
        self.global_rows = None  # T*B of the GLOBAL batch (BatchNorm count); None = local

    @property
    def device(self):
        return next(self.parameters()).device

    def initial_state(self, batch_size):
        return tuple(torch.zeros(self.core.num_layers, batch_size, self.core.hidden_size) for _ in range(2))

    # ------------------------------------------------------------------------------------------ parameters
    def _linears(self):
        off = 1 if self.batch_norm else 0
        return self.fc[off], self.fc[off + 2]

    def _param_list(self):
        """Order of the gradients returned by the backward."""
        l1, l2 = self._linears()
        ps = []
        if self.batch_norm:
            ps += [self.fc[0].weight, self.fc[0].bias]
        ps += [l1.weight, l1.bias, l2.weight, l2.bias]
        for l in range(2):
            ps += [getattr(self.core, f"weight_ih_l{l}"), getattr(self.core, f"weight_hh_l{l}"),
                   getattr(self.core, f"bias_ih_l{l}"), getattr(self.core, f"bias_hh_l{l}")]
        ps += [self.policy.weight, self.policy.bias]
        return ps

    def _refresh_weights(self):
        """bf16 (and transposed bf16) copies of the fp32 master weights; called at the start of every forward."""
        dev, bf = self.device, torch.bfloat16
        lib = _lib.lib()
        l1, l2 = self._linears()
        H, D, Dp = 1024, self.obs_size, _r64(self.obs_size)
        if self._wb is None or self._wb["dev"] != dev:
            z = lambda *s: torch.zeros(*s, dtype=bf, device=dev)  # noqa: E731
            self._wb = dict(dev=dev, W1=z(H, Dp), W1T=z(Dp, H), W2=z(H, H), W2T=z(H, H),
                            Wih=[z(4 * H, H), z(4 * H, H)], WihT=[z(H, 4 * H), z(H, 4 * H)],
                            Whh=[z(4 * H, H), z(4 * H, H)], WhhT=[z(H, 4 * H), z(H, 4 * H)])
        w = self._wb

        def cast(src, dst, dstT):
            r, c = src.shape
            _lib.check(lib.pvr_cast_weight(src.data_ptr(), r, c, dst.data_ptr(), dst.stride(0),
                                           dstT.data_ptr() if dstT is not None else None,
                                           dstT.stride(0) if dstT is not None else 0, _stream()), "pvr_cast_weight")

        cast(l1.weight.data, w["W1"], w["W1T"] if (self.batch_norm or self._needs_input_grad) else None)
        cast(l2.weight.data, w["W2"], w["W2T"])
        for l in range(2):
            cast(getattr(self.core, f"weight_ih_l{l}").data, w["Wih"][l], w["WihT"][l])
            cast(getattr(self.core, f"weight_hh_l{l}").data, w["Whh"][l], w["WhhT"][l])
        w["b_l"] = [getattr(self.core, f"bias_ih_l{l}").data + getattr(self.core, f"bias_hh_l{l}").data
                    for l in range(2)]

    def _lstm_chunks(self, T, B=None):
        """Number of time chunks of the two-layer wavefront (1 = layer after layer on one stream). With the persistent
        recurrence kernels (csrc/lstm_persist.cu: one launch per layer and direction, B <= 128) always 1. Otherwise 8
        while the step is being captured into a CUDA graph (launches cost no host time at replay), 1 in eager mode,
        where the step is bound by the host's launch rate and more launches would only slow it down."""
        env = os.environ.get("PVR_LSTM_CHUNKS")
        if not env and B is not None and _lib.lib().pvr_lstm_persist_supported(T, B, 1024):
            # The persistent kernels occupy 32 CTAs per batch tile of 32 rows: at B <= 32 (finetuning: B = 16) two of
            # them fit side by side, so the layers run as a wavefront over `c` chunks, each chunk one persistent launch
            # (state handed over through h_last / c_all / dh_rec / dc_rec, see _lstm_fwd_chunk / _lstm_bwd_chunk).
            # At B = 128 a layer fills 128 SMs and chunking only adds launches.
            c = int(os.environ.get("PVR_LSTM_PERSIST_CHUNKS", "2")) if B <= 32 else 1
            while c > 1 and (T % c or T // c < 8 or not _lib.lib().pvr_lstm_persist_supported(T // c, B, 1024)):
                c //= 2
            self._persist_chunks = True
            return max(c, 1)
        self._persist_chunks = False
        c = int(env) if env else (8 if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing() else 1)
        while c > 1 and (T % c or T // c < 2):
            c //= 2
        return max(c, 1)

    def _side_stream(self):
        if self._side is None or self._side.device != self.device:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _lstm_fwd_chunk(self, ws, w, l, t0, Tc, flags):
        B, H, r0 = ws.B, ws.H, t0 * ws.B
        h0 = ws.h0[l]
        if getattr(self, "_persist_chunks", False):
            # chunk of a persistent-kernel wavefront: an ordinary sequence (flags 0) whose initial state is the previous
            # chunk's final one — h_{t0-1} in h_last (fp32, masked and rounded to the bf16 operand exactly as the
            # unchunked kernel does between steps), c_{t0-1} already in c_all[t0]
            flags = 0
            if t0 > 0:
                h0 = ws.h_last[l]
        L = pvr_lstm_fwd(T=Tc, B=B, H=H, flags=flags, w_hh=w["Whh"][l].data_ptr(), xp=ws.XP[l][r0:].data_ptr(),
                         nd=ws.nd[t0:].data_ptr(), h0=h0.data_ptr(), c_all=ws.c_all[l][r0:].data_ptr(),
                         hm=ws.hm[l][r0:].data_ptr(), h_out=ws.HL[l][r0:].data_ptr(),
                         gates=ws.gates[l][r0:].data_ptr(), g_tmp=ws.g_tmp[l].data_ptr(),
                         h_last=ws.h_last[l].data_ptr(), counters=ws.lstm_counters[l].data_ptr(),
                         counters_bytes=ws.lstm_counters[l].numel() * 4)
        _lib.check(_lib.lib().pvr_lstm_forward(ctypes.byref(L), _stream()), "pvr_lstm_forward")

    def _lstm_bwd_chunk(self, ws, w, l, t0, Tc, flags, dbias=None):
        B, H, r0 = ws.B, ws.H, t0 * ws.B
        persist = getattr(self, "_persist_chunks", False)
        if persist:
            flags = 0  # dh_rec / dc_rec carry the state between the chunks (read on entry, dc_rec written on exit)
        L = pvr_lstm_bwd(T=Tc, B=B, H=H, flags=flags, w_hh_t=w["WhhT"][l].data_ptr(), nd=ws.nd[t0:].data_ptr(),
                         gates=ws.gates[l][r0:].data_ptr(), c_all=ws.c_all[l][r0:].data_ptr(),
                         dh_out=ws.dHL[l][r0:].data_ptr(), dh_rec=ws.dh_rec[l].data_ptr(),
                         dc_rec=ws.dc_rec[l].data_ptr(), dG=ws.dG[l][r0:].data_ptr(),
                         dbias=dbias.data_ptr() if dbias is not None else None,
                         counters=ws.lstm_counters[l].data_ptr(), counters_bytes=ws.lstm_counters[l].numel() * 4)
        _lib.check(_lib.lib().pvr_lstm_backward(ctypes.byref(L), _stream()), "pvr_lstm_backward")
        if persist and t0 > 0:
            # gradient flowing into h_{t0-1}: nd[t0] * (dG_{t0} W_hh) — the product the kernel leaves out at the first
            # step of its sequence (training starts from a constant state); same bf16 operands / fp32 sum as in-kernel
            gemm(ws.dG[l][r0:r0 + B], w["WhhT"][l], ws.dh_rec[l], B, H, 4 * H, out_f32=1)
            ws.dh_rec[l].mul_(ws.nd[t0].unsqueeze(1))

    def _check_generation(self, generation):
        if generation != self._generation:
            raise RuntimeError(
                "PolicyNet.backward: another forward of this module ran between this loss's forward and its backward "
                "(evaluation pass, second micro-batch, gradient accumulation over two forwards). The saved activations "
                "are kept in one workspace per module and have been overwritten — run backward() right after the "
                "forward it belongs to, or use a second PolicyNet instance (e.g. a `test_model`, as main_bc_2.py:100 "
                "does) for the interleaved passes.")

    def _workspace(self, T, B):
        key = (T, B, str(self.device))
        if key not in self._ws:
            if len(self._ws) > 4:
                # (evicted workspaces stay alive while a captured graph still points into them: BCTrainer and
                # _rollout_step hold references to the workspace they captured)
                self._ws.clear()
            self._ws[key] = _Workspace(T, B, self.obs_size, 1024, self.num_actions, self.device, self.batch_norm)
        return self._ws[key]

    # ------------------------------------------------------------------------------------------ forward (CUDA)
    def _forward_cuda(self, x, notdone, h0, c0):
        lib = _lib.lib()
        T, B = notdone.shape
        ws = self._workspace(T, B)
        M, H, D = ws.M, ws.H, ws.D
        self._generation += 1
        self._refresh_weights()
        w = self._wb
        l1, l2 = self._linears()
        ws.nd.copy_(notdone)
        x = x.contiguous()
        if self.batch_norm:
            bn = self.fc[0]
            if self.training:
                _lib.check(lib.pvr_bn1d_stats(x.data_ptr(), x.stride(0), M, D, ws.sums.data_ptr(), _stream()),
                           "pvr_bn1d_stats")
                count = float(M)
                if self.comm is not None:
                    self.comm.all_reduce(ws.sums)
                    count = float(self.global_rows)
                _lib.check(lib.pvr_bn1d_normalize(x.data_ptr(), x.stride(0), M, D, ws.sums.data_ptr(), count, bn.eps,
                                                  bn.momentum, bn.weight.data_ptr(), bn.bias.data_ptr(),
                                                  bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                                  ws.mean.data_ptr(), ws.rstd.data_ptr(), ws.X0.data_ptr(),
                                                  ws.X0.stride(0), _stream()), "pvr_bn1d_normalize")
                bn.num_batches_tracked += 1
            else:
                _lib.check(lib.pvr_bn1d_eval(x.data_ptr(), x.stride(0), M, D, bn.running_mean.data_ptr(),
                                             bn.running_var.data_ptr(), bn.eps, bn.weight.data_ptr(),
                                             bn.bias.data_ptr(), ws.mean.data_ptr(), ws.rstd.data_ptr(),
                                             ws.X0.data_ptr(), ws.X0.stride(0), _stream()), "pvr_bn1d_eval")
        else:
            _lib.check(lib.pvr_cast_rows_bf16(x.data_ptr(), x.stride(0), M, D, ws.X0.data_ptr(), ws.X0.stride(0),
                                              _stream()), "pvr_cast_rows_bf16")
        gemm(ws.X0, w["W1"], ws.H1, M, H, ws.Dp, bias=l1.bias.data, relu=True)
        gemm(ws.H1, w["W2"], ws.H2, M, H, H, bias=l2.bias.data, relu=True)
        # LSTM layers as a wavefront over time chunks: layer 0 works through chunk c + 1 on the current stream while
        # layer 1 (input projection of the chunk + recurrence) works through chunk c on a side stream. Every step is a
        # latency-bound GEMM + cell pair that fills a fraction of the GPU, so the two recurrences overlap.
        C = self._lstm_chunks(T, B)
        Tc = T // C
        cur = torch.cuda.current_stream(self.device)
        side = self._side_stream() if C > 1 else cur
        gemm(ws.H2, w["Wih"][0], ws.XP[0], M, 4 * H, H, bias=w["b_l"][0], out_f32=1)
        for l in range(2):
            ws.c_all[l][:B].copy_(c0[l])
            ws.h0[l].copy_(h0[l])
        for c in range(C):
            flags = (_lib.PVR_LSTM_CONT_PREV if c > 0 else 0) | (_lib.PVR_LSTM_CONT_NEXT if c < C - 1 else 0)
            r0, rows = c * Tc * B, Tc * B
            self._lstm_fwd_chunk(ws, w, 0, c * Tc, Tc, flags)
            if C > 1:
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
            with torch.cuda.stream(side):
                gemm(ws.HL[0][r0:r0 + rows], w["Wih"][1], ws.XP[1][r0:r0 + rows], rows, 4 * H, H, bias=w["b_l"][1],
                     out_f32=1)
                self._lstm_fwd_chunk(ws, w, 1, c * Tc, Tc, flags)
        if C > 1:
            ev = torch.cuda.Event()
            ev.record(side)
            cur.wait_event(ev)
        hn = [ws.h_last[l].clone() for l in range(2)]
        cn = [ws.c_all[l][M:M + B].clone() for l in range(2)]
        _lib.check(lib.pvr_heads_forward(ws.HL[1].data_ptr(), M, H, self.policy.weight.data_ptr(),
                                         self.policy.bias.data_ptr(), self.baseline.weight.data_ptr(),
                                         self.baseline.bias.data_ptr(), self.num_actions, ws.logits.data_ptr(),
                                         ws.baseline.data_ptr(), _stream()), "pvr_heads_forward")
        self._saved = (ws, x)
        return ws.logits.clone(), ws.baseline.clone(), torch.stack(hn), torch.stack(cn)

    # ------------------------------------------------------------------------------------------ backward (CUDA)
    def _backward_cuda(self, dlogits, need_dx=False):
        """Returns the parameter gradients (order of `_param_list`); with `need_dx` also d(loss)/d(input rows) as an
        fp32 (M, D) tensor (end-to-end finetuning: the input is the conv trunk's feature matrix)."""
        lib = _lib.lib()
        ws, x = self._saved
        w = self._wb
        T, B, M, Mp, H, D, Dp, A = ws.T, ws.B, ws.M, ws.Mp, ws.H, ws.D, ws.Dp, ws.A
        dev = self.device
        params = self._param_list()
        # one flat, zero-initialised gradient buffer (a single all-reduce under data parallelism); every view starts on
        # a 256-byte boundary so the fp32 TMA stores of the weight-gradient GEMMs are aligned
        sizes = [(p.numel() + 63) // 64 * 64 for p in params]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for p, sz in zip(params, sizes):
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += sz
        # gradient buckets of the data-parallel all-reduce, in the order the backward completes them: [LSTM layer 1 +
        # heads] after layer 1's weight gradients, [LSTM layer 0] after layer 0's, the trunk at the end. A bucket's
        # all-reduce runs on the communication stream while the GEMMs of the next bucket are computed.
        first_lstm = 6 if self.batch_norm else 4          # index of weight_ih_l0 in `params`
        off_l0, off_l1 = sum(sizes[:first_lstm]), sum(sizes[:first_lstm + 4])
        bucket_end = off_l0
        overlap = self.comm is not None and not need_dx

        def bucket(lo, hi):
            if overlap:
                self.comm.all_reduce(flat[lo:hi], wait=False)

        g = dict(zip(["bn_w", "bn_b"] if self.batch_norm else [], grads[:2]))
        names = ["W1", "b1", "W2", "b2", "Wih0", "Whh0", "bih0", "bhh0", "Wih1", "Whh1", "bih1", "bhh1", "Wp", "bp"]
        g.update(zip(names, grads[2 if self.batch_norm else 0:]))

        def colsum(src, n, out):
            _lib.check(lib.pvr_colsum_bf16(src.data_ptr(), src.stride(0), M, n, out.data_ptr(), _stream()),
                       "pvr_colsum_bf16")

        # heads: dHL1 = dlogits Wp, dWp, dbp
        _lib.check(lib.pvr_heads_backward(dlogits.data_ptr(), ws.HL[1].data_ptr(), self.policy.weight.data_ptr(), M, H,
                                          A, 1.0, ws.dHL[1].data_ptr(), g["Wp"].data_ptr(), g["bp"].data_ptr(),
                                          _stream()), "pvr_heads_backward")
        # reverse wavefront: layer 1 runs backwards through chunk c (+ the input gradient of that chunk for layer 0) on
        # the current stream while layer 0 runs backwards through chunk c + 1 on the side stream
        C = self._lstm_chunks(T, B)
        Tc = T // C
        cur = torch.cuda.current_stream(dev)
        side = self._side_stream() if C > 1 else cur
        for l in range(2):
            ws.dh_rec[l].zero_()
            ws.dc_rec[l].zero_()
        for c in reversed(range(C)):
            flags = (_lib.PVR_LSTM_CONT_PREV if c > 0 else 0) | (_lib.PVR_LSTM_CONT_NEXT if c < C - 1 else 0)
            r0, rows = c * Tc * B, Tc * B
            self._lstm_bwd_chunk(ws, w, 1, c * Tc, Tc, flags, g["bih1"])  # (+ the layer's bias gradient)
            # gradient w.r.t. layer-0 outputs (fp32, consumed by the layer-0 cell backward)
            gemm(ws.dG[1][r0:r0 + rows], w["WihT"][1], ws.dHL[0][r0:r0 + rows], rows, H, 4 * H, out_f32=1)
            if C > 1:
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
            with torch.cuda.stream(side):
                self._lstm_bwd_chunk(ws, w, 0, c * Tc, Tc, flags, g["bih0"])
        below = [ws.H2, ws.HL[0]]  # input of LSTM layer l
        for l in (1, 0):
            if l == 0 and C > 1:  # layer 1's weight gradients above overlap the tail of layer 0's recurrence
                ev = torch.cuda.Event()
                ev.record(side)
                cur.wait_event(ev)
            dG = ws.dG[l]
            g[f"bhh{l}"].copy_(g[f"bih{l}"])  # (accumulated by pvr_lstm_backward; both biases enter as their sum)
            # dW = dG^T X with dG (M, 4H) and X (M, H) as they sit in memory: MN-major tensor-core operands
            gemm(dG, ws.hm[l], g[f"Whh{l}"], 4 * H, H, M, out_f32=1, n_pad=H, mn=True)
            gemm(dG, below[l], g[f"Wih{l}"], 4 * H, H, M, out_f32=1, n_pad=H, mn=True)
            if l == 1:
                bucket(off_l1, flat.numel())
            else:
                bucket(off_l0, off_l1)
        # through ReLU of fc2: dZ2 = (dG0 W_ih0) * (H2 > 0)
        gemm(ws.dG[0], w["WihT"][0], ws.dZ2, M, H, 4 * H, res=ws.H2, res_mode=1)
        colsum(ws.dZ2, H, g["b2"])
        gemm(ws.dZ2, ws.H1, g["W2"], H, H, M, out_f32=1, n_pad=H, mn=True)
        gemm(ws.dZ2, w["W2T"], ws.dZ1, M, H, H, res=ws.H1, res_mode=1)
        colsum(ws.dZ1, H, g["b1"])
        if Dp == D:
            gemm(ws.dZ1, ws.X0, g["W1"], H, D, M, out_f32=1, n_pad=Dp, mn=True)
        else:
            gemm(ws.dZ1, ws.X0, ws.dW1p, H, Dp, M, out_f32=1, n_pad=Dp, mn=True)
            g["W1"].copy_(ws.dW1p[:, :D])
        dx = None
        if self.batch_norm or need_dx:
            if ws.dX0 is None:
                ws.dX0 = torch.zeros(M, Dp, dtype=torch.bfloat16, device=dev)
            gemm(ws.dZ1, w["W1T"], ws.dX0, M, Dp, H)
        if self.batch_norm:
            _lib.check(lib.pvr_bn1d_backward(ws.dX0.data_ptr(), ws.dX0.stride(0), x.data_ptr(), x.stride(0), M, D,
                                             ws.mean.data_ptr(), ws.rstd.data_ptr(), g["bn_w"].data_ptr(),
                                             g["bn_b"].data_ptr(), _stream()), "pvr_bn1d_backward")
        if need_dx:
            dx = torch.empty(M, D, dtype=torch.float32, device=dev)
            if self.batch_norm:
                count = float(M)
                sums = torch.stack([g["bn_w"], g["bn_b"]])  # sum dy*xhat, sum dy of THIS rank
                if self.comm is not None:
                    self.comm.all_reduce(sums)
                    count = float(self.global_rows)
                bn = self.fc[0]
                _lib.check(lib.pvr_bn1d_backward_dx(ws.dX0.data_ptr(), ws.dX0.stride(0), x.data_ptr(), x.stride(0), M,
                                                    D, ws.mean.data_ptr(), ws.rstd.data_ptr(), bn.weight.data_ptr(),
                                                    sums[0].data_ptr(), sums[1].data_ptr(), count, dx.data_ptr(), D,
                                                    _stream()), "pvr_bn1d_backward_dx")
            else:
                _lib.check(lib.pvr_bf16_rows_to_f32(ws.dX0.data_ptr(), ws.dX0.stride(0), M, D, dx.data_ptr(), D,
                                                    _stream()), "pvr_bf16_rows_to_f32")
        self._pending_flat = flat
        if not need_dx and self.comm is not None:
            # data parallel: SUM over ranks (the loss is pre-scaled by 1/global rows). The two LSTM buckets went out
            # while the backward continued (see `bucket` above); what is left is the trunk (BN, fc1, fc2).
            self.comm.all_reduce(flat[:bucket_end], wait=False)
            self.comm.join()
        return (grads, dx) if need_dx else grads

    # ------------------------------------------------------------------------------------------ public forward
    accepts_sample_action = True

    def forward(self, inputs, core_state=(), sample_action=True):
        """src/models.py:59-82. `sample_action=False` (not in the reference) skips the multinomial draw of the training
        mode, whose result the BC loops never read (main_bc_2.py:206-214 uses `policy_logits` only): ~15 small launches
        per step; the output's `action` is then None and the CUDA RNG is not advanced."""
        x = inputs['obs']  # (unroll_length, batch_size, obs_size)
        T, B, *_ = x.shape
        dev = self.device
        if dev.type != 'cuda':
            raise _lib.PvrError("PolicyNet: CUDA device required — pvr_habitat_b200 has no CPU fallback")
        x = torch.flatten(x, 0, 1).float().to(device=dev)
        notdone = (1 - inputs['done'].float()).abs().to(device=dev)
        if len(core_state) == 0:
            core_state = self.initial_state(B)
        h0, c0 = (s.to(device=dev, dtype=torch.float32) for s in core_state)
        params = self._param_list()
        action = None
        with torch.cuda.device(dev):
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                logits, baseline, hn, cn = _PolicyFn.apply(self, x, notdone, h0, c0, *params)
            elif not self.training and T * B <= self.rollout_graph_rows:
                logits, baseline, hn, cn, action = self._rollout_step(x, notdone, h0, c0)
            else:
                logits, baseline, hn, cn = self._forward_cuda(x, notdone, h0, c0)
        if self.training:
            action = torch.multinomial(F.softmax(logits, dim=1), num_samples=1) if sample_action else None
        elif action is None:
            action = torch.argmax(logits, dim=1)
        return dict(policy_logits=logits.view(T, B, -1), baseline=baseline.view(T, B),
                    action=action.view(T, B) if action is not None else None), (hn, cn)

    # ------------------------------------------------------------------------------------------ online rollout
    def _rollout_step(self, x, notdone, h0, c0):
        """Evaluation steps of the online rollout (src/test_model.py:11-17: T = B = 1, no grad, eval mode): the ~20
        launches of one step (weight casts included, so optimiser updates between rollouts are picked up) are captured
        once per (T, B) into a CUDA graph and replayed from fixed input / output buffers. The first call of a shape
        runs eagerly (lazy allocations, kernel attributes); results are the eager path's, bit for bit."""
        T, B = notdone.shape
        key = (T, B, str(self.device))
        if key not in self._rollout and len(self._rollout) >= 8:
            self._rollout.clear()  # many distinct shapes: drop the captured graphs (and the workspaces they pin)
        g = self._rollout.setdefault(key, {"calls": 0})
        g["calls"] += 1
        if g["calls"] == 1:
            logits, baseline, hn, cn = self._forward_cuda(x, notdone, h0, c0)
            return logits, baseline, hn, cn, torch.argmax(logits, dim=1)
        if "graph" not in g:
            g["in"] = tuple(t.clone() for t in (x, notdone, h0, c0))
            g["ws"] = self._workspace(T, B)  # keeps the captured buffers alive if the workspace cache is recycled
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward_cuda(*g["in"])
                g["out"] = out + (torch.argmax(out[0], dim=1),)
            g["graph"] = graph
        else:
            for dst, src in zip(g["in"], (x, notdone, h0, c0)):
                dst.copy_(src, non_blocking=True)
        g["graph"].replay()
        return tuple(t.clone() for t in g["out"])


class _CELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, inv_count):
        m, a = logits.shape
        loss = torch.zeros((), dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().pvr_ce_loss(logits.data_ptr(), targets.data_ptr(), m, a, inv_count, loss.data_ptr(),
                                              dl.data_ptr(), _stream()), "pvr_ce_loss")
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, grad):
        (dl,) = ctx.saved_tensors
        return dl * grad, None, None


def bc_loss(policy_logits, actions, global_rows=None):
    """F.nll_loss(F.log_softmax(flatten(policy_logits), -1), flatten(actions).long()) of main_bc_2.py:211-214 as one
    kernel (warp-shuffle reduction); `global_rows` = T*B of the global batch under data parallelism (loss and
    gradients are then pre-scaled so that a SUM all-reduce over ranks gives the reference's mean)."""
    logits = torch.flatten(policy_logits, 0, 1).contiguous().float()
    targets = torch.flatten(actions, 0, 1).contiguous().long()
    rows = logits.shape[0] if global_rows is None else global_rows
    return _CELossFn.apply(logits, targets, 1.0 / rows)


# ================================================================================================ finetuning
class _ConvPolicyFn(torch.autograd.Function):
    """PolicyNetWithConv as one autograd node: conv trunk (tcgen05 program) -> feature gather -> policy trunk."""

    @staticmethod
    def forward(ctx, net, obs_u8, notdone, h0, c0, *params):
        ctx.set_materialize_grads(False)
        out = net._forward_conv_cuda(obs_u8, notdone, h0, c0)
        ctx.net = net
        ctx.generation = net._generation
        ctx.n_params = len(params)
        ctx.mark_non_differentiable(out[2], out[3])
        return out

    @staticmethod
    def backward(ctx, dlogits, dbaseline, dh, dc):
        if dbaseline is not None:
            raise NotImplementedError("PolicyNetWithConv backward: a loss on `baseline` is not part of the BC path")
        if dlogits is None:
            return (None,) * (5 + ctx.n_params)
        ctx.net._check_generation(ctx.generation)
        return (None, None, None, None, None) + tuple(ctx.net._backward_conv_cuda(dlogits.contiguous()))


class PolicyNetWithConv(PolicyNet):
    """src/models.py:96-197: 5 x [Conv2d(3x3, stride 2, padding 1) + ELU] on every 3-channel frame of the uint8
    observation (divided by 255, H and W swapped by `transpose(1, 3)`), features of the frames concatenated along the
    last spatial axis, then the same trunk as PolicyNet. Trained end to end (main_bc_finetune.py).

    The forward of the conv trunk is the tcgen05 program of the 'random' PVR (program.add_small_conv, weights with
    their two spatial axes swapped instead of transposing the frames); the backward is a chain of tcgen05 GEMMs:
    dW = dZ^T col (split-K over the pixels) and dcol = dZ W followed by col2im."""

    def __init__(self, observation_shape, num_actions, batch_norm=False):
        nn.Module.__init__(self)
        in_channels = 3
        n_frames = observation_shape[2] // in_channels
        init_ = lambda m: init(m, nn.init.orthogonal_,  # noqa: E731
                               lambda x: nn.init.constant_(x, 0), nn.init.calculate_gain('relu'))
        layers = []
        for i in range(5):  # same construction order as src/models.py:107-118
            layers += [init_(nn.Conv2d(in_channels if i == 0 else 32, 32, kernel_size=(3, 3), stride=2, padding=1)),
                       nn.ELU()]
        self.feat_extract = nn.Sequential(*layers)
        H, W = observation_shape[0], observation_shape[1]
        if H != W or H % 32:
            raise NotImplementedError("PolicyNetWithConv: square frames with a side divisible by 32 (Habitat: 64x64)")
        self.frame_hw, self.n_frames = H, n_frames
        self.conv_hw = H // 32
        conv_out_size = 32 * self.conv_hw * self.conv_hw
        self._build_trunk(conv_out_size * n_frames, num_actions, batch_norm)
        self._needs_input_grad = True
        self._conv = None
        # PVR_SMALL_CONV_GEMM=1: the first conv layer (forward and weight gradient) stays on the implicit-GEMM path
        self._direct_first_layer = os.environ.get("PVR_SMALL_CONV_GEMM", "0") in ("", "0")

    def _conv_params(self):
        ps = []
        for i in (0, 2, 4, 6, 8):
            ps += [self.feat_extract[i].weight, self.feat_extract[i].bias]
        return ps

    def _forward_conv_cuda(self, obs_u8, notdone, h0, c0):
        from . import program as prg
        from .embeddings import Transforms
        lib = _lib.lib()
        dev = self.device
        TB, H, W, CN = obs_u8.shape
        N = self.n_frames
        F = TB * N
        st = self._conv if self._conv is not None and self._conv.get("F") == F else None
        if st is None:
            st = self._conv = dict(F=F)
            hw, sizes = H, []
            for _ in range(5):
                hw = (hw + 2 - 3) // 2 + 1
                sizes.append(hw)
            st["sizes"] = sizes
            st["tf"] = Transforms([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], size=H, crop=H)  # x / 255 only
        # The (tiny) program is compiled once per batch shape with every activation kept for the backward. The weights
        # change every step: their packed bf16 copies are rewritten in place ON THE DEVICE (no host round trip, no
        # synchronisation), H and W swapped instead of transposing the frames (src/models.py:169).
        if "enc" not in st:
            sd = {}
            for i in (0, 2, 4, 6, 8):
                sd[f"{i}.weight"] = self.feat_extract[i].weight.detach().float().cpu().transpose(2, 3).contiguous()
                sd[f"{i}.bias"] = self.feat_extract[i].bias.detach().float().cpu()
            prog = prg.Program()
            in_slot = prog.new_slot(H * H * 4)
            prog.emb_width = prg.add_small_conv(prog, sd, in_slot, 0, hw=H, keep_activations=True)
            enc = prog.finish(dev)
            enc.bind(F)
            st["enc"], st["slots"], st["emb_width"] = enc, [in_slot] + prog.kept_slots, prog.emb_width
            st["conv_ops"] = [i for i, m in enumerate(enc.op_meta) if m["kind"] == 1]
        enc = st["enc"]
        with torch.no_grad():
            for j, (i, op) in enumerate(zip((0, 2, 4, 6, 8), st["conv_ops"])):
                w = self.feat_extract[i].weight.detach().float().transpose(2, 3)
                packed = prg.pack_first_small_conv(w, 32) if j == 0 else prg.pack_small_conv(w, 32)
                t = enc.op_tensors[op]
                t["weight"][:packed.shape[0]].copy_(packed)
                t["bias"][:32].copy_(self.feat_extract[i].bias.detach().float())
        st["tf"].run(obs_u8, N, enc.slot0, _lib.PVR_FMT_NHWC4_BF16, True)
        if st.get("emb") is None or st["emb"].shape[0] != F:
            st["emb"] = torch.empty(F, st["emb_width"], dtype=torch.float32, device=dev)
        enc.forward(st["emb"], st["emb_width"])
        hc = self.conv_hw
        feat = torch.empty(TB, 32 * hc * hc * N, dtype=torch.float32, device=dev)
        _lib.check(lib.pvr_convfeat_gather(enc.slot_ptr(st["slots"][5]), 32, TB, N, hc, hc, 32, feat.data_ptr(),
                                           _stream()), "pvr_convfeat_gather")
        return self._forward_cuda(feat, notdone, h0, c0)

    def _backward_conv_cuda(self, dlogits):
        lib = _lib.lib()
        dev = self.device
        grads_policy, dx = self._backward_cuda(dlogits, need_dx=True)
        st = self._conv
        enc, slots, sizes, F = st["enc"], st["slots"], st["sizes"], st["F"]
        N, H, hc = self.n_frames, self.frame_hw, self.conv_hw
        TB = F // N
        bf, f32 = torch.bfloat16, torch.float32
        conv_params = self._conv_params()
        gconv = [torch.zeros_like(p, dtype=f32) for p in conv_params]
        # gradient w.r.t. the last conv output, back in NHWC frame order
        dy = torch.empty(F * hc * hc, 32, dtype=f32, device=dev)
        _lib.check(lib.pvr_convfeat_scatter(dx.data_ptr(), dx.stride(0), TB, N, hc, hc, 32, dy.data_ptr(), _stream()),
                   "pvr_convfeat_scatter")
        in_hw = [H] + sizes[:-1]
        # per-layer buffers of the backward, allocated once per batch shape: the padding of dzt (columns >= M) and of
        # colt (tap rows >= 9 Ci, columns >= M) is zeroed here and never written afterwards
        bufs = st.get("bwd_bufs")
        if bufs is None:
            bufs = st["bwd_bufs"] = []
            for l in range(5):
                M = F * sizes[l] * sizes[l]
                Mp = (M + 511) // 512 * 512
                Kp = 64 if l == 0 else 320
                if l == 0 and self._direct_first_layer:
                    bufs.append(dict(dw_nat=torch.zeros(32, 3, 3, 4, dtype=f32, device=dev)))
                    continue
                bufs.append(dict(dz=torch.empty(M, 64, dtype=bf, device=dev), dzt=torch.zeros(64, Mp, dtype=bf, device=dev),
                                 colt=torch.zeros(Kp, Mp, dtype=bf, device=dev), dw=torch.zeros(64, Kp, dtype=f32, device=dev),
                                 wt=torch.zeros(Kp, 64, dtype=bf, device=dev) if l > 0 else None,
                                 dcol=torch.empty(M, Kp, dtype=bf, device=dev) if l > 0 else None,
                                 dy_in=torch.empty(F * in_hw[l] * in_hw[l], 32, dtype=f32, device=dev) if l > 0 else None))
        for l in range(4, -1, -1):
            ho, hi = sizes[l], in_hw[l]
            ci = 4 if l == 0 else 32
            M = F * ho * ho
            Mp = (M + 511) // 512 * 512
            Kp = 64 if l == 0 else 320
            b = bufs[l]
            y_ptr, a_ptr = enc.slot_ptr(slots[l + 1]), enc.slot_ptr(slots[l])
            if l == 0 and self._direct_first_layer:
                # first layer (3 input channels, no input gradient needed): one mma.sync pass over dy / y / the frames
                # instead of ELU backward + im2col^T + a K = 3.3 M GEMM (csrc/small_conv.cu)
                dwn = b["dw_nat"]
                dwn.zero_()
                _lib.check(lib.pvr_small_conv1_wgrad(dy.data_ptr(), y_ptr, 32, a_ptr, F, hi, hi, ho, ho, dwn.data_ptr(),
                                                     gconv[1].data_ptr(), _stream()), "pvr_small_conv1_wgrad")
                gconv[0].copy_(dwn[..., :conv_params[0].shape[1]].permute(0, 3, 2, 1))
                continue
            # ELU backward + bias gradient + the transposed copy for the weight-gradient GEMM, one pass over dy
            dz, dzt, colt, dw = b["dz"], b["dzt"], b["colt"], b["dw"]
            _lib.check(lib.pvr_elu_backward_fused(dy.data_ptr(), y_ptr, 32, M, 32, dz.data_ptr(), dzt.data_ptr(), Mp,
                                                  gconv[2 * l + 1].data_ptr(), _stream()), "pvr_elu_backward_fused")
            _lib.check(lib.pvr_im2col_t(a_ptr, ci, F, hi, hi, ci, ho, ho, Mp, colt.data_ptr(), _stream()),
                       "pvr_im2col_t")
            chunks = Mp // 64
            split = 1
            while split < 64 and chunks % (split * 2) == 0:
                split *= 2
            dw.zero_()  # rows 32..63 unused (dZ^T padding); split-K slices accumulate into it
            gemm(dzt, colt, dw, 64, Kp, Mp, out_f32=2, split_k=split, n_pad=Kp)
            # natural layout (co, a, b, ci) -> parameter layout (co, ci, b, a): the conv runs on un-transposed frames
            w_nat = dw[:32, :9 * ci].view(32, 3, 3, ci)[..., :conv_params[2 * l].shape[1]]
            gconv[2 * l].copy_(w_nat.permute(0, 3, 2, 1))
            if l > 0:
                wt = b["wt"]  # (k = (a, b, ci), co); rows >= 288 and columns >= 32 stay zero
                w_t = conv_params[2 * l].detach().transpose(2, 3).permute(2, 3, 1, 0).reshape(9 * 32, 32)
                wt[:288, :32] = w_t.to(bf)
                dcol = b["dcol"]
                gemm(dz, wt, dcol, M, Kp, 64, n_pad=Kp)
                dy = b["dy_in"]
                _lib.check(lib.pvr_col2im(dcol.data_ptr(), Kp, F, hi, hi, 32, ho, ho, dy.data_ptr(), _stream()),
                           "pvr_col2im")
        if self.comm is not None:
            self.comm.all_reduce(self._pending_flat, wait=False)
            flat_c = torch.cat([g.flatten() for g in gconv])
            self.comm.all_reduce(flat_c)
            off = 0
            for g in gconv:
                g.copy_(flat_c[off:off + g.numel()].view_as(g))
                off += g.numel()
        return list(gconv) + list(grads_policy)

    # ------------------------------------------------------------------------------------------ public forward
    accepts_sample_action = True

    def forward(self, inputs, core_state=(), sample_action=True):
        x = inputs['obs']  # (unroll_length, batch_size, H, W, 3 * n_frames) uint8
        T, B, *_, CN = x.shape
        dev = self.device
        if dev.type != 'cuda':
            raise _lib.PvrError("PolicyNetWithConv: CUDA device required — pvr_habitat_b200 has no CPU fallback")
        if x.dtype != torch.uint8:
            raise _lib.PvrError("PolicyNetWithConv expects the uint8 frames the reference feeds it")
        obs = torch.flatten(x, 0, 1).to(device=dev).contiguous()
        notdone = (1 - inputs['done'].float()).abs().to(device=dev)
        if len(core_state) == 0:
            core_state = self.initial_state(B)
        h0, c0 = (s.to(device=dev, dtype=torch.float32) for s in core_state)
        params = self._conv_params() + self._param_list()
        with torch.cuda.device(dev):
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                logits, baseline, hn, cn = _ConvPolicyFn.apply(self, obs, notdone, h0, c0, *params)
            else:
                logits, baseline, hn, cn = self._forward_conv_cuda(obs, notdone, h0, c0)
        if self.training:
            action = torch.multinomial(F.softmax(logits, dim=1), num_samples=1) if sample_action else None
        else:
            action = torch.argmax(logits, dim=1)
        return dict(policy_logits=logits.view(T, B, -1), baseline=baseline.view(T, B),
                    action=action.view(T, B) if action is not None else None), (hn, cn)
