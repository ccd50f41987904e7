"""-m gpu: the CUDA BC-policy path (PolicyNet fwd/bwd, loss, fused optimizer, training loop) against the oracle and
the fixtures generated from the unmodified reference (tests/golden/policy.npz)."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate_policy as rp
from pvr_habitat_b200 import _lib, models
from pvr_habitat_b200.bc import BCTrainer
from pvr_habitat_b200.models import PolicyNet, bc_loss
from pvr_habitat_b200.optim import FusedAdam, FusedRMSprop

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "policy.npz"))


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ------------------------------------------------------------------------------------------------ GEMM variants
def test_gemm_fp32_output_with_bias():
    g = torch.Generator().manual_seed(0)
    m, n, k = 300, 4096, 1024
    a = torch.randn(m, k, generator=g).bfloat16().cuda()
    b = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().cuda()
    bias = torch.randn(n, generator=g).cuda()
    out = torch.full((m, n), float("nan"), device="cuda")
    models.gemm(a, b, out, m, n, k, bias=bias, out_f32=1)
    ref = a.float() @ b.float().t() + bias
    assert rel(out, ref) < 1e-5  # fp32 accumulate, fp32 out: only summation order differs


def test_gemm_split_k_atomic_accumulate():
    g = torch.Generator().manual_seed(1)
    m, n, k = 128, 1024, 4096
    a = torch.randn(m, k, generator=g).bfloat16().cuda()
    b = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().cuda()
    base = torch.randn(m, n, generator=g).cuda()
    out = base.clone()
    models.gemm(a, b, out, m, n, k, out_f32=2, split_k=8)
    assert rel(out, base + a.float() @ b.float().t()) < 1e-5


def test_gemm_relu_backward_mask():
    g = torch.Generator().manual_seed(2)
    m, n, k = 1000, 1024, 1024
    a = torch.randn(m, k, generator=g).bfloat16().cuda()
    b = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().cuda()
    act = torch.randn(m, n, generator=g).relu().bfloat16().cuda()
    out = torch.zeros(m, n, dtype=torch.bfloat16, device="cuda")
    models.gemm(a, b, out, m, n, k, res=act, res_mode=1)
    ref = (a.float() @ b.float().t()) * (act.float() > 0)
    assert rel(out.float(), ref) < 3e-3
    assert torch.all(out[act == 0] == 0)


# ------------------------------------------------------------------------------------------------ small kernels
def test_ce_loss_matches_torch():
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(777, 3, generator=g) * 3).cuda().requires_grad_(True)
    tgt = torch.randint(0, 3, (777,), generator=g).cuda()
    loss = bc_loss(logits.view(777, 1, 3), tgt.view(777, 1))
    loss.backward()
    l2 = logits.detach().clone().requires_grad_(True)
    ref = F.nll_loss(F.log_softmax(l2, -1), tgt)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    torch.testing.assert_close(logits.grad, l2.grad, atol=1e-8, rtol=1e-5)


@pytest.mark.parametrize("opt", ["rmsprop", "adam"])
def test_fused_optimizer_matches_torch(opt):
    g = torch.Generator().manual_seed(4)
    shapes = [(1024, 64), (1024,), (4096, 1024), (3, 1024), (3,)]
    p_ref = [torch.randn(*s, generator=g).cuda().requires_grad_(True) for s in shapes]
    p_new = [p.detach().clone().requires_grad_(True) for p in p_ref]
    if opt == "rmsprop":
        o_ref = torch.optim.RMSprop(p_ref, lr=1e-3, alpha=0.99, eps=1e-5)
        o_new = FusedRMSprop(p_new, lr=1e-3, alpha=0.99, eps=1e-5, max_grad_norm=40.0)
    else:
        o_ref = torch.optim.Adam(p_ref, lr=1e-3, eps=1e-8)
        o_new = FusedAdam(p_new, lr=1e-3, eps=1e-8, max_grad_norm=40.0)
    sched = [torch.optim.lr_scheduler.LambdaLR(o, lambda e: 1 - e / 10) for o in (o_ref, o_new)]
    for it in range(4):
        scale = 100.0 if it % 2 == 0 else 0.01  # clipping active / inactive
        grads = [torch.randn(*s, generator=g).cuda() * scale for s in shapes]
        for p, q, gr in zip(p_ref, p_new, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        for s in sched:
            s.step()
        norm = torch.nn.utils.clip_grad_norm_(p_ref, 40.0)
        o_ref.step()
        o_new.step()
        assert abs(float(o_new.gradient_norm()) - float(norm)) <= 1e-5 * float(norm)
        for p, q in zip(p_ref, p_new):
            torch.testing.assert_close(q, p, atol=2e-6, rtol=2e-5)
            torch.testing.assert_close(q.grad, p.grad, atol=1e-6, rtol=1e-5)  # clipped gradients stay visible


# ------------------------------------------------------------------------------------------------ PolicyNet
def make_policy(d, seed, batch_norm=True):
    torch.manual_seed(seed)
    net = PolicyNet((d,), 3, batch_norm)
    return net.cuda()


def test_policy_constructor_reproduces_reference_init(gold):
    net = make_policy(64, 7)
    sd = net.state_dict()
    sums = np.array([float(sd[str(n)].double().sum()) for n in gold["fb_param_names"]])
    # orthogonal_ runs a LAPACK QR: the last bits depend on the host CPU / thread count
    np.testing.assert_allclose(sums, gold["fb_param_sums"], rtol=1e-5, atol=1e-4)


def test_policy_forward_backward_vs_reference_golden(gold):
    """bf16 GEMM operands, fp32 accumulation: logits within 1e-2 relative of the fp32 reference module, gradient
    norms within 2 %; baseline.* receive no gradient, like the reference."""
    net = make_policy(64, 7)
    net.train()
    obs, done = torch.from_numpy(gold["fb_obs"]), torch.from_numpy(gold["fb_done"])
    state = (torch.from_numpy(gold["fb_h0"]), torch.from_numpy(gold["fb_c0"]))
    out, (hn, cn) = net(dict(obs=obs, done=done), state)
    assert out["policy_logits"].shape == (6, 5, 3) and out["baseline"].shape == (6, 5)
    assert out["action"].shape == (6, 5) and out["action"].dtype == torch.int64
    assert rel(out["policy_logits"], torch.from_numpy(gold["fb_logits"])) < 1e-2
    assert rel(out["baseline"], torch.from_numpy(gold["fb_baseline"])) < 2e-2
    assert rel(hn, torch.from_numpy(gold["fb_hn"])) < 1e-2 and rel(cn, torch.from_numpy(gold["fb_cn"])) < 1e-2
    torch.testing.assert_close(net.fc[0].running_var.cpu(), torch.from_numpy(gold["fb_running_var"]), rtol=1e-4,
                               atol=1e-6)
    loss = bc_loss(out["policy_logits"], torch.from_numpy(gold["fb_act"]).cuda())
    assert abs(float(loss) - float(gold["fb_loss"])) < 1e-2 * float(gold["fb_loss"])
    loss.backward()
    named = dict(net.named_parameters())
    for name, ref_norm in zip(gold["fb_param_names"], gold["fb_grad_norms"]):
        g = named[str(name)].grad
        if ref_norm < 0:
            assert g is None
        else:
            assert abs(float(g.norm()) - ref_norm) <= 0.02 * ref_norm + 1e-7, (name, float(g.norm()), ref_norm)
    assert rel(net.core.bias_ih_l1.grad, torch.from_numpy(gold["fb_grad_bias_ih_l1"])) < 2e-2
    assert rel(net.policy.weight.grad, torch.from_numpy(gold["fb_grad_policy_w"])) < 2e-2
    # trunk gradients: see test_policy_vs_oracle_other_shapes for where their ~6 % comes from
    assert rel(net.fc[0].weight.grad, torch.from_numpy(gold["fb_grad_bn_w"])) < 0.08


def test_policy_eval_mode_argmax_and_no_grad(gold):
    net = make_policy(64, 7).eval()
    obs, done = torch.from_numpy(gold["fb_obs"]), torch.from_numpy(gold["fb_done"])
    with torch.no_grad():
        out, _ = net(dict(obs=obs, done=done), net.initial_state(5))
    assert torch.equal(out["action"], out["policy_logits"].argmax(-1))


@pytest.mark.parametrize("T,B,D,bn", [(16, 32, 2048, True), (100, 16, 192, False), (7, 3, 130, True)])
def test_policy_vs_oracle_other_shapes(T, B, D, bn):
    """Production shape class (T=100, B=16), config-5 width (D=2048) and a ragged shape (rows and D not multiples
    of 64) against the CPU oracle."""
    net = make_policy(D, 11, bn).train()
    rng = np.random.default_rng(T * 1000 + B)
    obs = torch.from_numpy(rng.standard_normal((T, B, D)).astype(np.float32))
    done = torch.from_numpy(rng.random((T, B)) < 0.05)
    act = torch.from_numpy(rng.integers(0, 3, (T, B)))
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in net.state_dict().items()}
    out, _ = net(dict(obs=obs, done=done), net.initial_state(B))
    loss = bc_loss(out["policy_logits"], act.cuda())
    loss.backward()
    zero = (torch.zeros(2, B, 1024), torch.zeros(2, B, 1024))
    logits, _, _ = rp.policy_forward(sd, obs, done, zero, bn, True)
    ref_loss = rp.bc_loss(logits, act)
    ref_loss.backward()
    assert rel(out["policy_logits"], logits.detach()) < 1e-2
    assert abs(float(loss) - float(ref_loss)) < 1e-2 * float(ref_loss)
    agree = (out["policy_logits"].argmax(-1).cpu() == logits.argmax(-1)).float().mean()
    assert agree >= 0.98  # near-tied random-init logits: report, see DESIGN.md
    for k, p in net.named_parameters():
        if k.startswith("baseline."):
            continue
        # Measured (profiles/r02_grad_error_table.txt): LSTM / head gradients within 0.65 %, the BN / fc trunk's
        # within 6.2 %. A CPU emulation that rounds ONE class of tensors at a time (tools/grad_budget_cpu.py,
        # profiles/r02_grad_budget_cpu.txt) reproduces these numbers to three digits and attributes them: bf16 weights
        # alone 5 %, bf16 forward activations alone 4.8 % on the trunk (0.4 % on the LSTM) — the gradient of the bf16
        # FUNCTION differs from the fp32 one's, amplified through the recurrence on the way down to the trunk — while
        # the backward roundings (dG, dZ, dX0 in bf16) cost 0.2 % each. No backward precision would change that.
        assert rel(p.grad, sd[k].grad) < (0.08 if k.startswith("fc.") else 0.012), k


def test_bc_training_trace_vs_unmodified_reference(gold):
    """North star: BC loss curves within 1 % of the reference (main_bc_2.run trace, 12 steps, T=8, B=4)."""
    T, B, steps, seed = int(gold["bc_T"]), int(gold["bc_B"]), int(gold["bc_steps"]), int(gold["bc_seed"])
    torch.manual_seed(seed)
    random.seed(seed)
    net = PolicyNet((gold["bc_obs"].shape[1],), 3, True).cuda().train()
    tr = BCTrainer(net, gold["bc_obs"], gold["bc_action"], gold["bc_done"], B, T, int(gold["bc_max_frames"]))
    losses, norms = [], []
    for _ in range(steps):
        losses.append(float(tr.step()))
        norms.append(float(tr.gradient_norm()))
    np.testing.assert_allclose(losses, gold["bc_loss"], rtol=1e-2)
    np.testing.assert_allclose(norms, gold["bc_grad_norm"], rtol=5e-2)


def test_bc_host_batches_equal_device_resident(gold):
    """The reference's host gather + H2D data path and the device-resident gather give identical steps."""
    T, B = 8, 4
    out = []
    for host in (False, True):
        torch.manual_seed(3)
        random.seed(3)
        net = PolicyNet((gold["bc_obs"].shape[1],), 3, True).cuda().train()
        tr = BCTrainer(net, gold["bc_obs"], gold["bc_action"], gold["bc_done"], B, T, 4 * T * B, host_batches=host)
        out.append([float(tr.step()) for _ in range(3)])
    # fp32 atomics (split-K, column sums) make the last bits run-dependent
    np.testing.assert_allclose(out[0], out[1], rtol=1e-4)


# ------------------------------------------------------------------------------------------------ finetuning (a13)
def test_policy_with_conv_forward_backward_vs_reference_golden(gold):
    """PolicyNetWithConv (src/models.py:96-197): conv trunk forward on tcgen05, backward as GEMMs (dW = dZ^T col,
    dcol = dZ W + col2im), through BatchNorm's input gradient and the LSTM policy."""
    from pvr_habitat_b200.models import PolicyNetWithConv
    torch.manual_seed(13)
    net = PolicyNetWithConv((64, 64, 6), 3, batch_norm=True).cuda().train()
    sd = net.state_dict()
    sums = np.array([float(sd[str(n)].double().sum()) for n in gold["cv_param_names"]])
    np.testing.assert_allclose(sums, gold["cv_param_sums"], rtol=1e-5, atol=1e-4)
    obs, done = torch.from_numpy(gold["cv_obs"]), torch.from_numpy(gold["cv_done"])
    out, _ = net(dict(obs=obs, done=done), net.initial_state(obs.shape[1]))
    # 12 rows only: BatchNorm over so few rows amplifies the bf16 rounding of the conv features, and the logits of the
    # fresh network are ~1e-2 in magnitude; the loss (criterion) is checked to 1 %
    assert rel(out["policy_logits"], torch.from_numpy(gold["cv_logits"])) < 3e-2
    loss = bc_loss(out["policy_logits"], torch.from_numpy(gold["cv_act"]).cuda())
    assert abs(float(loss) - float(gold["cv_loss"])) < 1e-2 * float(gold["cv_loss"])
    loss.backward()
    named = dict(net.named_parameters())
    for name, ref_norm in zip(gold["cv_param_names"], gold["cv_grad_norms"]):
        g = named[str(name)].grad
        if ref_norm < 0:
            assert g is None
        else:
            assert abs(float(g.norm()) - ref_norm) <= 0.1 * ref_norm + 1e-7, (name, float(g.norm()), ref_norm)
    assert rel(net.feat_extract[0].weight.grad, torch.from_numpy(gold["cv_grad_conv0_w"])) < 0.15
    # element-wise agreement of weight-gradient tensors: bf16 GEMM operands + a 12-row BatchNorm (cancelling sums)
    assert rel(net.feat_extract[8].weight.grad, torch.from_numpy(gold["cv_grad_conv4_w"])) < 0.2
    assert rel(net.feat_extract[4].bias.grad, torch.from_numpy(gold["cv_grad_conv2_b"])) < 0.2


def test_finetune_loss_curve_vs_oracle():
    """main_bc_finetune-style training (RMSprop, clip 40, LambdaLR) of PolicyNetWithConv for 8 steps against the
    oracle restatement on the same data / seeds: loss within 1 %."""
    from oracle import restate
    from pvr_habitat_b200.models import PolicyNetWithConv
    from pvr_habitat_b200.optim import FusedRMSprop
    T, B, steps = 4, 4, 8
    frames = restate.structured_frames(256, 64, 64, 6, 71)
    rng = np.random.default_rng(8)
    action = rng.integers(0, 3, 256)
    done = rng.random(256) < 0.03
    # ---- CUDA path
    torch.manual_seed(21)
    random.seed(21)
    net = PolicyNetWithConv((64, 64, 6), 3, batch_norm=True).cuda().train()
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    opt = FusedRMSprop(net.parameters(), lr=1e-4, alpha=0.99, eps=1e-5, max_grad_norm=40.0)
    max_epochs = steps + 1
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 1 - e / max_epochs)
    from pvr_habitat_b200.utils_bc import sample_with_minimum_distance, window_indices
    got, starts_log = [], []
    for _ in range(steps):
        starts = sample_with_minimum_distance(256, B, T)
        starts_log.append(starts)
        idx = window_indices(starts, T, 256)
        o = torch.from_numpy(frames[idx])
        out, _ = net(dict(obs=o, done=torch.from_numpy(done[idx])), net.initial_state(B))
        loss = bc_loss(out["policy_logits"], torch.from_numpy(action[idx]).cuda())
        sched.step()
        opt.zero_grad()
        loss.backward()
        opt.step()
        got.append(float(loss))
    # ---- oracle (torch CPU autograd + torch RMSprop, the reference's own optimiser classes)
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items()
              if v.is_floating_point() and "running" not in k and not k.startswith("baseline.")}
    ropt = torch.optim.RMSprop(list(params.values()), lr=1e-4, alpha=0.99, eps=1e-5)
    rsched = torch.optim.lr_scheduler.LambdaLR(ropt, lambda e: 1 - e / max_epochs)
    ref = []
    full = dict(sd0)
    for starts in starts_log:
        idx = window_indices(starts, T, 256)
        full.update(params)
        zero = (torch.zeros(2, B, 1024), torch.zeros(2, B, 1024))
        logits, _, _ = rp.policy_conv_forward(full, torch.from_numpy(frames[idx]), torch.from_numpy(done[idx]), zero,
                                              True, True)
        loss = rp.bc_loss(logits, torch.from_numpy(action[idx]))
        rsched.step()
        ropt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 40.0)
        ropt.step()
        ref.append(float(loss))
    np.testing.assert_allclose(got, ref, rtol=1e-2)


# ------------------------------------------------------------------------------------------------ online rollout
def test_rollout_step_graph_replay_equals_eager():
    """src/test_model.py:11-17 (T = B = 1, eval, no grad, state carried over): the first step runs eagerly, the
    second is captured into a CUDA graph, later ones replay it — all equal to the plain forward, also after the
    weights were updated in place between steps (the graph re-casts the bf16 weight copies)."""
    net = make_policy(2048, 5).eval()
    plain = make_policy(2048, 5).eval()
    plain.rollout_graph_rows = 0
    with torch.no_grad():
        net.fc[0].running_mean.normal_(0, 0.1)
        net.fc[0].running_var.uniform_(0.5, 1.5)
    plain.load_state_dict(net.state_dict())
    s1 = tuple(s.cuda() for s in net.initial_state(1))
    s2 = tuple(s.clone() for s in s1)
    g = torch.Generator().manual_seed(0)
    for k in range(7):
        obs = torch.randn(1, 1, 2048, generator=g)
        done = torch.tensor([[k == 4]])
        if k == 5:  # an optimiser step between two rollouts: raw in-place update, as FusedRMSprop does
            with torch.no_grad():
                for a, b in zip(net.parameters(), plain.parameters()):
                    d = torch.randn(a.shape, generator=g).cuda() * 1e-2
                    a.add_(d)
                    b.add_(d)
        with torch.no_grad():
            o1, s1 = net(dict(obs=obs, done=done), s1)
            o2, s2 = plain(dict(obs=obs, done=done), s2)
        # same kernels on the same data; the recurrent GEMM accumulates its two split-K slices with fp32 atomics, so
        # the last bits depend on the arrival order (also between two eager runs)
        for key in ("policy_logits", "baseline"):
            assert rel(o1[key], o2[key]) < 1e-5, (k, key, o1[key], o2[key])
        assert torch.equal(o1["action"], o2["action"]), k
        assert o1["policy_logits"].shape == (1, 1, 3) and o1["action"].shape == (1, 1)
        assert all(rel(a, b) < 1e-5 for a, b in zip(s1, s2)), k
    assert "graph" in net._rollout[(1, 1, str(net.device))] and not plain._rollout
    # training-mode and large forwards never take the graph path
    net.train()
    with torch.no_grad():
        net(dict(obs=torch.randn(1, 2, 2048), done=torch.zeros(1, 2, dtype=torch.bool)), net.initial_state(2))
    net.eval()
    with torch.no_grad():
        net(dict(obs=torch.randn(3, 4, 2048), done=torch.zeros(3, 4, dtype=torch.bool)), net.initial_state(4))
    assert list(net._rollout) == [(1, 1, str(net.device))] and net._rollout[(1, 1, str(net.device))]["calls"] == 7


@pytest.mark.parametrize("T,B", [(64, 16), (100, 4), (16, 8)])
def test_lstm_wavefront_chunks_equal_the_layer_by_layer_schedule(T, B, monkeypatch):
    """The two implementations of the recurrence on the same data: one chunk per layer = the persistent whole-sequence
    kernels (csrc/lstm_persist.cu; deterministic sums, exp-based sigmoid / tanh with the fast exponential) against the
    per-step kernels run as a two-stream wavefront over time chunks (PVR_LSTM_CHUNKS = 8: recurrent state carried
    across chunk borders forwards and backwards, done masks at the borders; expf / tanhf, 8-slice atomic GEMMs in the
    backward). They share the rounding points (bf16 recurrent operand, bf16 dG), so the results agree to the noise of
    a bf16 rounding of h landing on the other side (~1e-3 relative on the logits)."""
    monkeypatch.setenv("PVR_LSTM_NO_INPLACE", "1")
    rng = np.random.default_rng(T + B)
    obs = torch.from_numpy(rng.standard_normal((T, B, 192)).astype(np.float32))
    done = torch.from_numpy(rng.random((T, B)) < 0.1)
    act = torch.from_numpy(rng.integers(0, 3, (T, B))).cuda()
    h0 = tuple(torch.from_numpy(rng.standard_normal((2, B, 1024)).astype(np.float32)) for _ in range(2))
    results = []
    for chunks in ("1", "8"):
        monkeypatch.setenv("PVR_LSTM_CHUNKS", chunks)
        net = make_policy(192, 3).train()
        assert net._lstm_chunks(T) == (1 if chunks == "1" else {64: 8, 100: 4, 16: 8}[T])
        out, state = net(dict(obs=obs, done=done), h0)
        bc_loss(out["policy_logits"], act).backward()
        torch.cuda.synchronize()
        results.append((out["policy_logits"].detach(), state, {k: p.grad.clone() for k, p in net.named_parameters()
                                                               if p.grad is not None}))
    (l1, s1, g1), (l2, s2, g2) = results
    assert rel(l1, l2) < 3e-3 and all(rel(a, b) < 3e-3 for a, b in zip(s1, s2)), (rel(l1, l2),)
    assert set(g1) == set(g2)
    for k in g1:
        assert rel(g1[k], g2[k]) < 2e-2, (k, rel(g1[k], g2[k]))  # dG is rounded to bf16 after the noisy fp32 sums


def test_trained_policy_argmax_agreement_vs_oracle():
    """North star: policy action argmax identical on >= 99.9 % of frames. A policy trained for 40 BC steps on the CUDA
    path (logit margins opened up), then the CUDA eval forward against the fp32 oracle on 4096 frames with the same
    weights."""
    D, n = 256, 8192
    rng = np.random.default_rng(0)
    obs = np.maximum(rng.standard_normal((n, D)).astype(np.float32), 0)
    w = rng.standard_normal((D, 3)).astype(np.float32) / np.sqrt(D)
    action = (obs @ w + 0.3 * rng.standard_normal((n, 3)).astype(np.float32)).argmax(1)
    done = rng.random(n) < 1 / 200
    torch.manual_seed(1)
    random.seed(1)
    net = PolicyNet((D,), 3, True).cuda().train()
    tr = BCTrainer(net, obs, action, done, 32, 16, 10 ** 9)
    losses = [float(tr.step()) for _ in range(40)]
    assert losses[-1] < losses[0]
    net.eval()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    o = torch.from_numpy(obs[:4096]).view(64, 64, D)
    d = torch.from_numpy(done[:4096]).view(64, 64)
    with torch.no_grad():
        out, _ = net(dict(obs=o, done=d), net.initial_state(64))
        ref, _, _ = rp.policy_forward(sd, o, d, (torch.zeros(2, 64, 1024), torch.zeros(2, 64, 1024)), True, False)
    agree = float((out["policy_logits"].argmax(-1).cpu() == ref.argmax(-1)).float().mean())
    assert torch.equal(out["action"], out["policy_logits"].argmax(-1))
    assert rel(out["policy_logits"], ref) < 1e-2 and agree >= 0.999, (agree, rel(out["policy_logits"], ref))
