"""Per-parameter relative L2 error of the CUDA policy's gradients against the fp32 oracle at a few shapes.
Usage: python tools/grad_error_table.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restate_policy as rp  # noqa: E402
from pvr_habitat_b200.models import PolicyNet, bc_loss  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


for T, B, D, bn in [(16, 32, 2048, True), (100, 16, 192, False), (7, 3, 130, True), (64, 128, 2048, True)]:
    torch.manual_seed(7)
    net = PolicyNet((D,), 3, batch_norm=bn).cuda().train()
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
    print(f"--- T={T} B={B} D={D} bn={bn}: logits rel {rel(out['policy_logits'], logits.detach()):.2e}, "
          f"loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    tot_n = tot_d = 0.0
    for k, p in net.named_parameters():
        if k.startswith("baseline."):
            continue
        r = rel(p.grad, sd[k].grad)
        tot_n += float((p.grad.double().cpu() - sd[k].grad.double()).pow(2).sum())
        tot_d += float(sd[k].grad.double().pow(2).sum())
        print(f"    {k:24s} rel-L2 {r:.4f}   |g| {float(sd[k].grad.norm()):.3e}")
    print(f"    whole gradient vector: rel-L2 {np.sqrt(tot_n / tot_d):.4f}")
