"""ORACLE — test infrastructure only. CPU restatement (torch fp32) of the reference's BC policy path:
PolicyNet forward (src/models.py:57-89), BC loss (main_bc_2.py:211-214), grad-norm statistic + clip + RMSprop with
LambdaLR (main_bc_2.py:80-90, 216-227) and the BC sampling loop (main_bc_2.py:186-204, src/utils_bc.py:24-29).

Pinned against the unmodified reference by oracle/make_golden.py -> tests/golden/policy.npz (PolicyNet outputs and
gradient norms; per-step loss / gradient-norm trace of main_bc_2.run on a synthetic pickle).
"""
import random

import numpy as np
import torch
import torch.nn.functional as F

H = 1024


def policy_forward(sd, obs, done, core_state, batch_norm, training=True, bn_buffers=None):
    """obs (T,B,D) float, done (T,B) bool, core_state = (h (2,B,H), c (2,B,H)); sd: reference state_dict keys.
    Returns policy_logits (T,B,A), baseline (T,B), (h, c). Differentiable w.r.t. the tensors in `sd`."""
    T, B = obs.shape[:2]
    x = torch.flatten(obs, 0, 1).float()
    off = 0
    if batch_norm:  # nn.BatchNorm1d in train mode: batch mean / biased variance (src/models.py:30-34)
        if training:
            mean = x.mean(0)
            var = x.var(0, unbiased=False)
            if bn_buffers is not None:  # running stats: momentum 0.1, unbiased variance
                n = x.shape[0]
                bn_buffers["running_mean"].mul_(0.9).add_(0.1 * mean.detach())
                bn_buffers["running_var"].mul_(0.9).add_(0.1 * var.detach() * n / max(n - 1, 1))
        else:
            mean, var = sd["fc.0.running_mean"], sd["fc.0.running_var"]
        x = (x - mean) / torch.sqrt(var + 1e-5) * sd["fc.0.weight"] + sd["fc.0.bias"]
        off = 1
    x = F.relu(F.linear(x, sd[f"fc.{off}.weight"], sd[f"fc.{off}.bias"]))
    x = F.relu(F.linear(x, sd[f"fc.{off + 2}.weight"], sd[f"fc.{off + 2}.bias"]))
    core_input = x.view(T, B, -1)
    notdone = (1 - done.float()).abs()
    h, c = [core_state[0][0], core_state[0][1]], [core_state[1][0], core_state[1][1]]
    outs = []
    for t in range(T):  # src/models.py:68-72: mask the state of BOTH layers, then one 2-layer LSTM step
        nd = notdone[t].view(-1, 1)
        inp = core_input[t]
        for l in range(2):
            hp, cp = nd * h[l], nd * c[l]
            g = F.linear(inp, sd[f"core.weight_ih_l{l}"], sd[f"core.bias_ih_l{l}"]) + \
                F.linear(hp, sd[f"core.weight_hh_l{l}"], sd[f"core.bias_hh_l{l}"])
            i, f, gg, o = g.chunk(4, 1)  # PyTorch gate order i, f, g, o
            c[l] = torch.sigmoid(f) * cp + torch.sigmoid(i) * torch.tanh(gg)
            h[l] = torch.sigmoid(o) * torch.tanh(c[l])
            inp = h[l]
        outs.append(inp)
    core_output = torch.cat(outs, 0)
    logits = F.linear(core_output, sd["policy.weight"], sd["policy.bias"])
    baseline = F.linear(core_output, sd["baseline.weight"], sd["baseline.bias"])
    return logits.view(T, B, -1), baseline.view(T, B), (torch.stack(h), torch.stack(c))


def bc_loss(logits, actions):
    """main_bc_2.py:211-214."""
    return F.nll_loss(F.log_softmax(torch.flatten(logits, 0, 1), dim=-1), torch.flatten(actions, 0, 1).long())


def sample_with_minimum_distance(n, k, d):
    """src/utils_bc.py:17-29."""
    sample = random.sample(range(n - (k - 1) * (d - 1)), k)
    order = sorted(range(k), key=lambda i: sample[i])
    ranks = sorted(order, key=lambda i: order[i])
    return [s + (d - 1) * r for s, r in zip(sample, ranks)]


def make_batch(obs, action, done, starting_i, T):
    """main_bc_2.py:191-201: windows of T consecutive samples wrapping modulo n."""
    n = len(action)
    idx = [np.mod(np.arange(i, i + T), n) for i in starting_i]
    o = np.stack([obs[ix] for ix in idx], axis=1)
    a = np.stack([action[ix] for ix in idx], axis=1)
    d = np.stack([done[ix] for ix in idx], axis=1)
    return o, a, d


def bc_train(sd, obs, action, done, T, B, steps, max_frames, batch_norm, lr=1e-4, alpha=0.99, eps=1e-5,
             max_grad_norm=40.0, seed=1, conv=False):
    """The training loop of main_bc_2.py:186-227 restated; `sd` tensors are updated in place.
    Returns per-step (loss, pre-clip gradient norm). conv=True: main_bc_finetune.py:167-208 (raw uint8 observations
    through the conv trunk of PolicyNetWithConv, trained end to end)."""
    random.seed(seed)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k}
    bn_buffers = {"running_mean": sd["fc.0.running_mean"].clone(), "running_var": sd["fc.0.running_var"].clone()} \
        if batch_norm else None
    square_avg = {k: torch.zeros_like(v) for k, v in params.items()}
    max_epochs = max_frames // (T * B) + 1
    trace = []
    for step in range(steps):
        starting_i = sample_with_minimum_distance(len(action), B, T)
        o, a, d = make_batch(obs, action, done, starting_i, T)
        state = (torch.zeros(2, B, H), torch.zeros(2, B, H))
        if conv:
            feat = conv_features(params, torch.from_numpy(o)).view(T, B, -1)
            logits, _, _ = policy_forward(params, feat, torch.from_numpy(d), state, batch_norm, True, bn_buffers)
        else:
            logits, _, _ = policy_forward(params, torch.from_numpy(o), torch.from_numpy(d), state, batch_norm, True,
                                          bn_buffers)
        loss = bc_loss(logits, torch.from_numpy(a))
        lr_k = lr * (1 - (step + 1) / max_epochs)  # scheduler.step() precedes optimizer.step() (main_bc_2.py:216)
        grads = torch.autograd.grad(loss, [params[k] for k in params], allow_unused=True)
        gmap = {k: g for k, g in zip(params, grads) if g is not None}  # baseline.* get no gradient
        norm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in gmap.values())))
        coef = min(1.0, max_grad_norm / (norm + 1e-6))  # torch.nn.utils.clip_grad_norm_
        with torch.no_grad():
            for k, g in gmap.items():
                g = g * coef
                square_avg[k].mul_(alpha).addcmul_(g, g, value=1 - alpha)
                params[k].addcdiv_(g, square_avg[k].sqrt().add_(eps), value=-lr_k)  # eps outside the sqrt
        trace.append((float(loss.detach()), norm))
    for k in params:
        sd[k] = params[k].detach()
    return trace


def bc_train_conv(sd, obs_u8, action, done, T, B, steps, max_frames, batch_norm=True, **kw):
    return bc_train(sd, obs_u8, action, done, T, B, steps, max_frames, batch_norm, conv=True, **kw)


def synthetic_bc_data(n, d, n_actions, seed):
    """Structured embeddings, actions from a hidden linear map (so the loss can fall), done ~ Bernoulli(0.02)."""
    rng = np.random.default_rng(seed)
    latent = rng.standard_normal((n, 16)).astype(np.float32)
    mix = rng.standard_normal((16, d)).astype(np.float32) / 4
    obs = (latent @ mix + 0.3 * rng.standard_normal((n, d)).astype(np.float32) + 0.5).astype(np.float32)
    wmap = rng.standard_normal((16, n_actions)).astype(np.float32)
    action = np.argmax(latent @ wmap + 0.5 * rng.standard_normal((n, n_actions)), axis=1).astype(np.int64)
    done = rng.random(n) < 0.02
    reward = np.zeros(n, dtype=np.float32)
    return obs, action, done, reward


def synthetic_frame_trajectories(n_traj, traj_len, seed, hw=64, n_frames=2):
    """Raw-frame trajectory pickle of behavioral_cloning/save_opt_trajectories.py:94-106 (lists over trajectories;
    obs (L, hw, hw, 3n) uint8) with structured frames; actions follow the mean colour of the central patch of the
    first frame plus noise, so the loss can fall. Returns (pickle dict, flat action array)."""
    from oracle import restate
    rng = np.random.default_rng(seed)
    frames = restate.structured_frames(n_traj * traj_len, hw, hw, 3 * n_frames, seed)
    q = hw // 4
    feat = frames[:, q:3 * q, q:3 * q, :3].reshape(len(frames), -1, 3).mean(1)
    action = np.argmax(feat + 8 * rng.standard_normal(feat.shape), 1).astype(np.int64)
    cut = lambda a: [a[i * traj_len:(i + 1) * traj_len] for i in range(n_traj)]  # noqa: E731
    data = dict(obs=cut(frames), action=cut(action),
                reward=[np.zeros(traj_len, np.float32) for _ in range(n_traj)],
                done=[np.arange(traj_len) == traj_len - 1 for _ in range(n_traj)],
                true_state=[np.zeros((traj_len, 12)) for _ in range(n_traj)])
    return data, action


def init_policy_state(obs_size, num_actions, batch_norm, seed, _rng_state=None):
    """Initial state_dict of the reference's PolicyNet((obs_size,), num_actions, batch_norm) built right after
    torch.manual_seed(seed): same layer order and init calls as src/models.py:17-44 (orthogonal, gain sqrt(2) for the
    trunk, 1 for the heads, zero biases, nn.LSTM default init), hence the same draws from the torch RNG."""
    from torch import nn
    if _rng_state is not None:
        torch.set_rng_state(_rng_state)  # continue an RNG stream (PolicyNetWithConv builds its conv trunk first)
    else:
        torch.manual_seed(seed)
    gain = nn.init.calculate_gain('relu')

    def make(i, o, g):  # construct, then re-initialise, one layer at a time (the order the RNG is consumed in)
        m = nn.Linear(i, o)
        nn.init.orthogonal_(m.weight.data, gain=g)
        nn.init.constant_(m.bias.data, 0)
        return m

    fc1 = make(obs_size, H, gain)
    fc2 = make(H, H, gain)
    bn = nn.BatchNorm1d(obs_size) if batch_norm else None
    core = nn.LSTM(H, H, 2)
    policy = make(H, num_actions, 1)
    baseline = make(H, 1, 1)
    sd, off = {}, 0
    if batch_norm:
        sd.update({"fc.0." + k: v.detach().clone() for k, v in bn.state_dict().items()})
        off = 1
    sd.update({f"fc.{off}." + k: v.detach().clone() for k, v in fc1.state_dict().items()})
    sd.update({f"fc.{off + 2}." + k: v.detach().clone() for k, v in fc2.state_dict().items()})
    sd.update({"core." + k: v.detach().clone() for k, v in core.state_dict().items()})
    sd.update({"policy." + k: v.detach().clone() for k, v in policy.state_dict().items()})
    sd.update({"baseline." + k: v.detach().clone() for k, v in baseline.state_dict().items()})
    return sd


def conv_features(sd, obs_u8):
    """PolicyNetWithConv feature path (src/models.py:159-171): obs (T,B,H,W,3n) uint8 -> (T*B, conv_out*n) float.
    x/255, split into 3-channel frames, `transpose(1, 3)` (channels first, H and W swapped), 5 x [conv3x3 s2 p1 + ELU],
    concatenation of the frames along the LAST spatial axis, flatten."""
    x = torch.flatten(obs_u8, 0, 1).float() / 255.
    feats = []
    for fr in torch.split(x, 3, -1):
        y = fr.transpose(1, 3)
        for i in (0, 2, 4, 6, 8):
            y = F.elu(F.conv2d(y, sd[f"feat_extract.{i}.weight"], sd[f"feat_extract.{i}.bias"], stride=2, padding=1))
        feats.append(y)
    x = torch.cat(feats, -1)
    return x.reshape(x.shape[0], -1)


def policy_conv_forward(sd, obs_u8, done, core_state, batch_norm, training=True):
    T, B = obs_u8.shape[:2]
    feat = conv_features(sd, obs_u8).view(T, B, -1)
    return policy_forward(sd, feat, done, core_state, batch_norm, training)


def init_policy_conv_state(frame_hw, n_frames, num_actions, batch_norm, seed):
    """Initial state_dict of the reference's PolicyNetWithConv((hw, hw, 3n), A, bn) right after manual_seed(seed):
    conv layers first (orthogonal, gain sqrt(2)), then the PolicyNet trunk (src/models.py:100-150)."""
    from torch import nn
    torch.manual_seed(seed)
    gain = nn.init.calculate_gain('relu')
    sd = {}
    for j, i in enumerate((0, 2, 4, 6, 8)):
        m = nn.Conv2d(3 if j == 0 else 32, 32, kernel_size=(3, 3), stride=2, padding=1)
        nn.init.orthogonal_(m.weight.data, gain=gain)
        nn.init.constant_(m.bias.data, 0)
        sd[f"feat_extract.{i}.weight"], sd[f"feat_extract.{i}.bias"] = m.weight.detach().clone(), m.bias.detach().clone()
    d = 32 * (frame_hw // 32) ** 2 * n_frames
    state = torch.get_rng_state()
    trunk = init_policy_state(d, num_actions, batch_norm, seed=None, _rng_state=state)
    sd.update(trunk)
    return sd
