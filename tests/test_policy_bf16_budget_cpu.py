"""Error budget of the CUDA policy's bf16 operands, on CPU: the fp32 oracle against a torch emulation that rounds to
bf16 exactly where libpvr_b200 does (BatchNorm output, weights, fc activations, the masked recurrent operand hm_t, the
LSTM layer outputs; fp32 accumulation, gates, cell state and heads). North star: "policy action argmax identical on
>= 99.9 % of frames". With random-init weights half of the frames have a top-2 logit margin below 1e-2 and bf16 cannot
keep every argmax (99.7 % here; the -m gpu tests assert >= 98 % there); after a few dozen BC steps the margins open up
and the budget is met. This is an emulation, not the CUDA path — it says what the design can reach; the CUDA kernels are
compared with the oracle in tests/test_gpu_policy.py."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import restate_policy as rp


def bf(t):
    return t.to(torch.bfloat16).float()


def forward_bf16_emulation(sd, obs, done, core_state):
    """Eval-mode forward with batch_norm=True, rounding points of models.PolicyNet._forward_cuda / csrc/policy.cu."""
    T, B = obs.shape[:2]
    x = torch.flatten(obs, 0, 1).float()
    x = (x - sd["fc.0.running_mean"]) / torch.sqrt(sd["fc.0.running_var"] + 1e-5) * sd["fc.0.weight"] + sd["fc.0.bias"]
    x = bf(x)                                                                         # X0
    x = bf(F.relu(F.linear(x, bf(sd["fc.1.weight"]), sd["fc.1.bias"])))               # H1
    x = bf(F.relu(F.linear(x, bf(sd["fc.3.weight"]), sd["fc.3.bias"])))               # H2
    nd = (1 - done.float()).abs()
    inp = x.view(T, B, -1)
    for l in range(2):
        bias = sd[f"core.bias_ih_l{l}"] + sd[f"core.bias_hh_l{l}"]
        xp = F.linear(inp.reshape(T * B, -1), bf(sd[f"core.weight_ih_l{l}"]), bias).view(T, B, -1)  # fp32 projection
        whh = bf(sd[f"core.weight_hh_l{l}"])
        h, c, outs = core_state[0][l], core_state[1][l], []
        for t in range(T):
            m = nd[t].view(-1, 1)
            i, f, g, o = (xp[t] + F.linear(bf(m * h), whh)).chunk(4, 1)               # hm_t is bf16
            c = torch.sigmoid(f) * (m * c) + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(bf(h))                                                        # layer output is bf16
        inp = torch.stack(outs)
    return F.linear(inp.reshape(T * B, -1), sd["policy.weight"], sd["policy.bias"]).view(T, B, -1)


def test_bf16_operands_keep_the_argmax_of_a_trained_policy():
    torch.set_num_threads(min(8, torch.get_num_threads()))
    D, n = 256, 8192
    rng = np.random.default_rng(0)
    obs = np.maximum(rng.standard_normal((n, D)).astype(np.float32), 0)
    w = rng.standard_normal((D, 3)).astype(np.float32) / np.sqrt(D)
    action = (obs @ w + 0.3 * rng.standard_normal((n, 3)).astype(np.float32)).argmax(1)
    done = rng.random(n) < 1 / 200
    sd = rp.init_policy_state(D, 3, True, 1)
    o = torch.from_numpy(obs[:4096]).view(64, 64, D)
    d = torch.from_numpy(done[:4096]).view(64, 64)
    zero = (torch.zeros(2, 64, 1024), torch.zeros(2, 64, 1024))

    def agreement():
        with torch.no_grad():
            ref, _, _ = rp.policy_forward(sd, o, d, zero, True, False)
            emu = forward_bf16_emulation(sd, o, d, zero)
        top2 = ref.topk(2, -1).values
        return (float((ref.argmax(-1) == emu.argmax(-1)).float().mean()), float((ref - emu).norm() / ref.norm()),
                float((top2[..., 0] - top2[..., 1]).median()))

    a0, rel0, margin0 = agreement()
    assert a0 >= 0.98 and rel0 <= 1e-2 and margin0 < 0.05          # random init: near-tied logits
    trace = rp.bc_train(sd, obs, action, done, 16, 32, 40, 10 ** 9, True)
    assert trace[-1][0] < trace[0][0]                               # the loss fell: margins opened up
    a1, rel1, margin1 = agreement()
    assert margin1 > 3 * margin0
    assert a1 >= 0.999 and rel1 <= 1e-2, (a1, rel1)                 # the north-star budget
