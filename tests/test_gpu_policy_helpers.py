"""Small policy kernels against torch on the same inputs: heads (src/models.py:56-57 policy / baseline linears), the
weight casts / transposes done after every optimiser step, and the global gradient norm of main_bc_2.py:220-224.
Both the vectorised kernels (aligned, multiples of 4 / 8) and the scalar fallbacks are covered."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from pvr_habitat_b200 import _lib as L
    return L, L.lib()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("m,K,A", [(8192, 1024, 3), (37, 1024, 8), (100, 192, 5), (64, 16384, 2)])
def test_heads_forward_backward(m, K, A):
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(m + K + A)
    h = torch.randn(m, K, device="cuda", generator=g).bfloat16()
    Wp = torch.randn(A, K, device="cuda", generator=g) * 0.05
    bp = torch.randn(A, device="cuda", generator=g)
    Wb = torch.randn(1, K, device="cuda", generator=g) * 0.05
    bb = torch.randn(1, device="cuda", generator=g)
    logits = torch.empty(m, A, device="cuda")
    base = torch.empty(m, device="cuda")
    L.check(lib.pvr_heads_forward(h.data_ptr(), m, K, Wp.data_ptr(), bp.data_ptr(), Wb.data_ptr(), bb.data_ptr(), A,
                                  logits.data_ptr(), base.data_ptr(), _stream()), "pvr_heads_forward")
    hf = h.double()
    ref_l = hf @ Wp.double().t() + bp.double()
    ref_b = (hf @ Wb.double().t()).squeeze(1) + bb.double()
    assert torch.allclose(logits.double(), ref_l, rtol=1e-5, atol=1e-4)
    assert torch.allclose(base.double(), ref_b, rtol=1e-5, atol=1e-4)

    dl = torch.randn(m, A, device="cuda", generator=g) / m
    dh = torch.empty(m, K, device="cuda")
    dWp = torch.zeros(A, K, device="cuda")
    dbp = torch.zeros(A, device="cuda")
    L.check(lib.pvr_heads_backward(dl.data_ptr(), h.data_ptr(), Wp.data_ptr(), m, K, A, 1.0, dh.data_ptr(),
                                   dWp.data_ptr(), dbp.data_ptr(), _stream()), "pvr_heads_backward")
    assert torch.allclose(dh.double(), dl.double() @ Wp.double(), rtol=1e-5, atol=1e-7)
    ref_dw = dl.double().t() @ hf
    assert torch.allclose(dWp.double(), ref_dw, rtol=1e-4, atol=1e-6 * float(ref_dw.abs().max()) + 1e-7)
    assert torch.allclose(dbp.double(), dl.double().sum(0), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("rows,cols", [(4096, 1024), (1024, 2048), (3, 1024), (100, 36), (130, 66)])
@pytest.mark.parametrize("transposed", [False, True])
def test_cast_weight(rows, cols, transposed):
    L, lib = _lib()
    w = torch.randn(rows, cols, device="cuda")
    ldb = (cols + 63) // 64 * 64
    ldt = (rows + 63) // 64 * 64
    wb = torch.zeros(rows, ldb, device="cuda", dtype=torch.bfloat16)
    wt = torch.zeros(cols, ldt, device="cuda", dtype=torch.bfloat16) if transposed else None
    L.check(lib.pvr_cast_weight(w.data_ptr(), rows, cols, wb.data_ptr(), ldb, wt.data_ptr() if transposed else None,
                                ldt if transposed else 0, _stream()), "pvr_cast_weight")
    ref = w.bfloat16()
    assert torch.equal(wb[:, :cols], ref)
    assert not wb[:, cols:].any()
    if transposed:
        assert torch.equal(wt[:, :rows], ref.t())
        assert not wt[:, rows:].any()


def test_sumsq_matches_double_sum():
    L, lib = _lib()
    sizes = [4096 * 1024, 1024, 3, 2048 * 1024 + 2, 7, 1 << 20]
    flat = torch.randn(sum(sizes) + 8, device="cuda")
    ts, off = [], 1  # offset 1: the second and later tensors are not 16-byte aligned in general
    for n in sizes:
        ts.append(flat[off:off + n])
        off += n
    VP = ctypes.c_void_p * len(ts)
    I64 = ctypes.c_int64 * len(ts)
    out = torch.zeros(1, device="cuda", dtype=torch.float64)
    L.check(lib.pvr_optim_sumsq(VP(*[t.data_ptr() for t in ts]), I64(*sizes), len(ts), out.data_ptr(), _stream()),
            "pvr_optim_sumsq")
    ref = sum(float((t.double() ** 2).sum()) for t in ts)
    assert abs(float(out) - ref) <= 1e-10 * ref


@pytest.mark.parametrize("F,H", [(6, 64), (3, 30), (2, 17)])
def test_small_conv1_wgrad_vs_torch(F, H):
    """First-layer weight / bias gradient of the small-conv trunk (src/models.py:108) against torch autograd on the
    same bf16-rounded operands: dz = bf16(dy * ELU'(y)), products accumulated in fp32."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(F * 100 + H)
    Ho = (H - 1) // 2 + 1
    x = torch.zeros(F, H, H, 4, device="cuda")
    x[..., :3] = torch.rand(F, H, H, 3, device="cuda", generator=g)
    xb = x.bfloat16()
    y = (torch.randn(F, Ho, Ho, 32, device="cuda", generator=g) * 0.7).bfloat16()  # ELU output: > -1
    y = torch.where(y.float() < -0.99, torch.full_like(y, -0.5), y)
    dy = torch.randn(F * Ho * Ho, 32, device="cuda", generator=g)
    dw = torch.zeros(32, 3, 3, 4, device="cuda")
    db = torch.zeros(32, device="cuda")
    L.check(lib.pvr_small_conv1_wgrad(dy.data_ptr(), y.data_ptr(), 32, xb.data_ptr(), F, H, H, Ho, Ho, dw.data_ptr(),
                                      db.data_ptr(), _stream()), "pvr_small_conv1_wgrad")
    yf = y.float().reshape(-1, 32)
    dz = (dy * torch.where(yf > 0, torch.ones_like(yf), yf + 1)).bfloat16().double()
    xin = xb.double().permute(0, 3, 1, 2)[:, :3].contiguous().requires_grad_(False)
    w = torch.zeros(32, 3, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
    out = torch.nn.functional.conv2d(xin, w, stride=2, padding=1)  # (F, 32, Ho, Ho)
    out.backward(dz.reshape(F, Ho, Ho, 32).permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1)  # (co, a, b, c)
    scale = float(ref.abs().max())
    assert float((dw[..., :3].double() - ref).abs().max()) < 2e-5 * scale + 1e-6
    assert not dw[..., 3].any()
    ref_b = dz.sum(0)
    assert float((db.double() - ref_b).abs().max()) < 2e-5 * float(ref_b.abs().max()) + 1e-6
