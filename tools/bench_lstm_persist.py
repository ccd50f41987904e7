"""Per-time-step latency of the persistent LSTM kernels (forward and backward) at a few batch sizes.
Usage: python tools/bench_lstm_persist.py [T]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_lstm_persist as tp  # noqa: E402
from pvr_habitat_b200 import _lib  # noqa: E402
from pvr_habitat_b200._lib import pvr_lstm_bwd, pvr_lstm_fwd  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = 1024
print("supported:", {B: _lib.lib().pvr_lstm_persist_supported(T, B, H) for B in (16, 32, 64, 128)})
for B in (16, 64, 128):
    w_hh, xp, nd, h0, c0 = tp._problem(T, B, 1)
    a = tp._forward_cuda(w_hh, xp, nd, h0, c0, T, B)
    L = pvr_lstm_fwd(T=T, B=B, H=H, flags=0, w_hh=w_hh.data_ptr(), xp=a["xp"].data_ptr(), nd=nd.data_ptr(),
                     h0=h0.data_ptr(), c_all=a["c_all"].data_ptr(), hm=a["hm"].data_ptr(), h_out=a["h_out"].data_ptr(),
                     gates=a["gates"].data_ptr(), g_tmp=a["g_tmp"].data_ptr(), h_last=a["h_last"].data_ptr())
    w_hh_t = w_hh.t().contiguous()
    dh_out = torch.randn(T * B, H, device="cuda") * 0.1
    dh_rec, dc_rec = torch.zeros(B, H, device="cuda"), torch.zeros(B, H, device="cuda")
    dG = torch.zeros(T * B, 4 * H, dtype=torch.bfloat16, device="cuda")
    Lb = pvr_lstm_bwd(T=T, B=B, H=H, flags=0, w_hh_t=w_hh_t.data_ptr(), nd=nd.data_ptr(), gates=a["gates"].data_ptr(),
                      c_all=a["c_all"].data_ptr(), dh_out=dh_out.data_ptr(), dh_rec=dh_rec.data_ptr(),
                      dc_rec=dc_rec.data_ptr(), dG=dG.data_ptr())
    for name, fn, desc in (("forward", _lib.lib().pvr_lstm_forward, L), ("backward", _lib.lib().pvr_lstm_backward, Lb)):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            _lib.check(fn(ctypes.byref(desc), _lib.current_stream_ptr()))
        ev[0].record()
        for _ in range(20):
            _lib.check(fn(ctypes.byref(desc), _lib.current_stream_ptr()))
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 20
        print(f"B={B:4d} T={T} {name:8s}: {ms * 1e3:8.1f} us per launch, {ms * 1e3 / T:6.2f} us per time step")

# ---- phase profile of the forward / backward kernel at B = 128 (clock64 stamps, one SM clock per CTA)
import numpy as np  # noqa: E402
NAMES = ["first ready", "all ready", "last TMA issued", "first chunk landed", "last chunk landed", "acc complete",
         "pieces sent", "exchange done", "cell done", "after bar", "-", "after atomic", "chunk 4", "chunk 8", "chunk 12"]
B = 128
w_hh, xp, nd, h0, c0 = tp._problem(T, B, 1)
a = tp._forward_cuda(w_hh, xp, nd, h0, c0, T, B)
prof = torch.zeros(128 * 64 * 16, dtype=torch.int64, device="cuda")
_lib.lib().pvr_lstm_persist_profile(ctypes.c_void_p(prof.data_ptr()))
for which in ("forward", "backward"):
    prof.zero_()
    if which == "forward":
        tp._forward_cuda(w_hh, xp, nd, h0, c0, T, B)
    else:
        tp._backward_cuda(w_hh, nd, a["gates"], a["c_all"], torch.randn(T * B, H, device="cuda") * 0.1, T, B)
    pr = prof.cpu().numpy().reshape(128, 64, 16).astype(np.float64)
    print(f"--- {which}: mean cycles since the previous step's 'after atomic' (steps 8..40), per CTA")
    for blk in (0, 5, 64, 127):
        rows = []
        for s in range(8, min(40, T - 1)):
            base = pr[blk, s - 1, 11]
            rows.append(pr[blk, s, :15] - base)
        m = np.mean(rows, 0)
        # slots a kernel does not stamp stay 0 and come out hugely negative: not printed
        print(f"block {blk:3d}: " + ", ".join(f"{n} {v:.0f}" for n, v in zip(NAMES, m) if n != "-" and abs(v) < 1e7))
_lib.lib().pvr_lstm_persist_profile(None)
