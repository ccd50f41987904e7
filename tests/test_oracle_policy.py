"""Pins oracle/restate_policy.py against the reference-generated fixture tests/golden/policy.npz (CPU only)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate_policy as rp


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "policy.npz"))


def test_initial_weights_match_reference_constructor(gold):
    sd = rp.init_policy_state(64, 3, True, 7)
    names = [str(n) for n in gold["fb_param_names"]]
    sums = np.array([float(sd[n].double().sum()) for n in names])
    # orthogonal_ runs a LAPACK QR: the last bits depend on the host CPU / thread count
    np.testing.assert_allclose(sums, gold["fb_param_sums"], rtol=1e-5, atol=1e-4)


def test_forward_backward_matches_reference_module(gold):
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in rp.init_policy_state(64, 3, True, 7).items()}
    buffers = {"running_mean": sd["fc.0.running_mean"].clone(), "running_var": sd["fc.0.running_var"].clone()}
    state = (torch.from_numpy(gold["fb_h0"]), torch.from_numpy(gold["fb_c0"]))
    logits, baseline, (hn, cn) = rp.policy_forward(sd, torch.from_numpy(gold["fb_obs"]),
                                                   torch.from_numpy(gold["fb_done"]), state, True, True, buffers)
    np.testing.assert_allclose(logits.detach().numpy(), gold["fb_logits"], atol=2e-6)
    np.testing.assert_allclose(baseline.detach().numpy(), gold["fb_baseline"], atol=2e-6)
    np.testing.assert_allclose(hn.detach().numpy(), gold["fb_hn"], atol=2e-6)
    np.testing.assert_allclose(cn.detach().numpy(), gold["fb_cn"], atol=2e-6)
    np.testing.assert_allclose(buffers["running_var"].numpy(), gold["fb_running_var"], rtol=1e-5)
    loss = rp.bc_loss(logits, torch.from_numpy(gold["fb_act"]))
    assert abs(float(loss) - float(gold["fb_loss"])) < 1e-6
    loss.backward()
    for name, ref_norm in zip(gold["fb_param_names"], gold["fb_grad_norms"]):
        g = sd[str(name)].grad
        if ref_norm < 0:  # the reference leaves baseline.* without gradient (unused in the BC loss)
            assert g is None and str(name).startswith("baseline.")
        else:
            assert abs(float(g.norm()) - ref_norm) <= 1e-4 * ref_norm + 1e-8, name
    np.testing.assert_allclose(sd["core.bias_ih_l1"].grad.numpy(), gold["fb_grad_bias_ih_l1"], atol=1e-7)
    np.testing.assert_allclose(sd["policy.weight"].grad.numpy(), gold["fb_grad_policy_w"], atol=1e-7)
    np.testing.assert_allclose(sd["fc.0.weight"].grad.numpy(), gold["fb_grad_bn_w"], atol=1e-7)


def test_bc_training_trace_matches_unmodified_main_bc_2(gold):
    """Loss / gradient-norm trace of the reference's main_bc_2.run() (12 steps, T=8, B=4, batch_norm)."""
    T, B, steps = int(gold["bc_T"]), int(gold["bc_B"]), int(gold["bc_steps"])
    seed = int(gold["bc_seed"])
    sd = rp.init_policy_state(gold["bc_obs"].shape[1], 3, True, seed)
    trace = rp.bc_train(sd, gold["bc_obs"], gold["bc_action"], gold["bc_done"], T, B, steps,
                        int(gold["bc_max_frames"]), True, seed=seed)
    loss = np.array([t[0] for t in trace])
    norm = np.array([t[1] for t in trace])
    np.testing.assert_allclose(loss, gold["bc_loss"], rtol=2e-4)
    np.testing.assert_allclose(norm, gold["bc_grad_norm"], rtol=2e-3)


def test_sampler_matches_reference_draws():
    import random
    from pvr_habitat_b200 import utils_bc
    for seed in range(4):
        random.seed(seed)
        a = rp.sample_with_minimum_distance(1000, 16, 50)
        random.seed(seed)
        b = utils_bc.sample_with_minimum_distance(1000, 16, 50)
        assert a == b
        assert min(np.diff(sorted(a))) >= 50
    idx = utils_bc.window_indices([998, 3], 5, 1000)
    assert idx.shape == (5, 2) and list(idx[:, 0]) == [998, 999, 0, 1, 2]


def test_policy_with_conv_restatement_matches_reference_module(gold):
    sd = rp.init_policy_conv_state(64, 2, 3, True, 13)
    names = [str(n) for n in gold["cv_param_names"]]
    sums = np.array([float(sd[n].double().sum()) for n in names])
    np.testing.assert_allclose(sums, gold["cv_param_sums"], rtol=1e-5, atol=1e-4)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in sd.items()}
    obs, done = torch.from_numpy(gold["cv_obs"]), torch.from_numpy(gold["cv_done"])
    B = obs.shape[1]
    zero = (torch.zeros(2, B, 1024), torch.zeros(2, B, 1024))
    logits, _, _ = rp.policy_conv_forward(sd, obs, done, zero, True, True)
    np.testing.assert_allclose(logits.detach().numpy(), gold["cv_logits"], atol=5e-6)
    loss = rp.bc_loss(logits, torch.from_numpy(gold["cv_act"]))
    assert abs(float(loss.detach()) - float(gold["cv_loss"])) < 1e-6
    loss.backward()
    for name, ref_norm in zip(names, gold["cv_grad_norms"]):
        g = sd[name].grad
        if ref_norm < 0:
            assert g is None
        else:
            assert abs(float(g.norm()) - ref_norm) <= 1e-3 * ref_norm + 1e-9, name
    np.testing.assert_allclose(sd["feat_extract.0.weight"].grad.numpy(), gold["cv_grad_conv0_w"], atol=1e-6, rtol=1e-3)
