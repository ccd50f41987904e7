import sys, os, time, random
sys.path.insert(0, '/root/repo')
import torch, bench
from pvr_habitat_b200.bc import BCTrainer
from pvr_habitat_b200.models import PolicyNetWithConv
obs, action, done = bench.finetune_dataset()
torch.manual_seed(1); random.seed(1)
net = PolicyNetWithConv((64, 64, 6), 3, batch_norm=True).cuda().train()
tr = BCTrainer(net, obs, action, done, 16, 100, 10 ** 9)
print("use_graph", tr.use_graph)
ls=[]
for _ in range(8): ls.append(float(tr.step()))
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(40): l = tr.step()
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print(f"{40/dt:.1f} steps/s {dt/40*1e3:.2f} ms/step; losses {ls[:3]} ... {float(l):.4f}")
