"""CPU emulation of the policy forward / backward with bf16 rounding switched on ONE class of tensors at a time (weights,
forward activations, dG, dZ, dX0): which rounding point the gradient deviation from the fp32 reference comes from.
Result (T=16, B=32, D=2048; profiles/r02_grad_budget_cpu.txt): the backward roundings cost 0.2 % each; rounding the
FORWARD activations alone moves the fc-trunk weight gradients by 3.5-5 % (LSTM 0.4 %): the gradient of the bf16
function differs from the gradient of the fp32 function, no backward precision changes that."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, torch.nn.functional as F
from oracle import restate_policy as rp
torch.set_num_threads(8)
def bf(t): return t.to(torch.bfloat16).float()
class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd): ctx.bwd = bwd; return bf(x) if fwd else x
    @staticmethod
    def backward(ctx, g): return (bf(g) if ctx.bwd else g), None, None
def R(x, fwd, bwd=False): return RoundSTE.apply(x, fwd, bwd)

def run(sd, obs, done, act, cfg):
    T,B = obs.shape[:2]
    p = {k: v.clone().requires_grad_(True) for k,v in sd.items() if v.is_floating_point() and 'running' not in k}
    W = lambda k: R(p[k], cfg['w'])   # bf16 weights in the GEMMs, gradient lands on the fp32 master copy
    x = torch.flatten(obs,0,1)
    mean = x.mean(0); var = x.var(0, unbiased=False)
    x = (x-mean)/torch.sqrt(var+1e-5)*p['fc.0.weight']+p['fc.0.bias']
    x = R(x, cfg['act'], cfg['dX0'])                      # X0 bf16; dX0 bf16
    z1 = F.linear(x, W('fc.1.weight'), p['fc.1.bias']); z1 = R(z1, False, cfg['dZ'])   # dZ1 rounded
    h1 = R(F.relu(z1), cfg['act'])
    z2 = F.linear(h1, W('fc.3.weight'), p['fc.3.bias']); z2 = R(z2, False, cfg['dZ'])   # dZ2 rounded
    h2 = R(F.relu(z2), cfg['act'])
    nd = (1-done.float()).abs()
    inp = h2.view(T,B,-1)
    for l in range(2):
        bias = p[f'core.bias_ih_l{l}']+p[f'core.bias_hh_l{l}']
        xp = F.linear(inp.reshape(T*B,-1), W(f'core.weight_ih_l{l}'), bias).view(T,B,-1)
        whh = W(f'core.weight_hh_l{l}')
        h = torch.zeros(B,1024); c = torch.zeros(B,1024); outs=[]
        for t in range(T):
            m = nd[t].view(-1,1)
            gates = xp[t] + F.linear(R(m*h, cfg['act']), whh)
            gates = R(gates, False, cfg['dG'])            # dG rounded to bf16
            i,f,g,o = gates.chunk(4,1)
            c = torch.sigmoid(f)*(m*c)+torch.sigmoid(i)*torch.tanh(g)
            h = torch.sigmoid(o)*torch.tanh(c)
            outs.append(R(h, cfg['act']))
        inp = torch.stack(outs)
    logits = F.linear(inp.reshape(T*B,-1), p['policy.weight'], p['policy.bias']).view(T,B,-1)
    loss = rp.bc_loss(logits, act)
    loss.backward()
    return {k: v.grad for k,v in p.items() if v.grad is not None}

T,B,D = 16,32,2048
sd = rp.init_policy_state(D,3,True,7)
rng = np.random.default_rng(T*1000+B)
obs = torch.from_numpy(rng.standard_normal((T,B,D)).astype(np.float32))
done = torch.from_numpy(rng.random((T,B))<0.05); act = torch.from_numpy(rng.integers(0,3,(T,B)))
base = run(sd,obs,done,act,dict(w=False,act=False,dG=False,dZ=False,dX0=False))
def rel(a,b): return float((a-b).norm()/b.norm())
keys=['fc.0.weight','fc.1.weight','fc.3.weight','core.weight_ih_l0','core.weight_hh_l1','policy.weight']
for name,cfg in [('weights only',dict(w=True,act=False,dG=False,dZ=False,dX0=False)),
                 ('fwd activations only',dict(w=False,act=True,dG=False,dZ=False,dX0=False)),
                 ('dG only',dict(w=False,act=False,dG=True,dZ=False,dX0=False)),
                 ('dZ only',dict(w=False,act=False,dG=False,dZ=True,dX0=False)),
                 ('dX0 only',dict(w=False,act=False,dG=False,dZ=False,dX0=True)),
                 ('all',dict(w=True,act=True,dG=True,dZ=True,dX0=True))]:
    g = run(sd,obs,done,act,cfg)
    print(f"{name:22s}", ' '.join(f"{k.split('.')[0][:4]+k[-9:]}:{rel(g[k],base[k]):.4f}" for k in keys))
