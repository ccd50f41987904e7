"""Is the embedding pass bound by kernel time or by gaps between its 91 launches? One default-workload pass (192 x 3 frames)
eager, replayed from a CUDA graph, and as the sum of per-op device times. Result (B200, round 2): 13.01 / 13.11 / 13.18 ms —
no gaps: programmatic dependent launch already hides the launch latency; the step is kernel time."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch, bench
net = bench.build_net("moco_aug_uber_34", torch.device("cuda", 0))
n_frames, obs_n = 3, 192
obs = [torch.from_numpy(bench.make_observations(obs_n, n_frames, 100 + i)).cuda() for i in range(3)]
out = torch.empty(obs_n, n_frames * net.out_size, device="cuda")
def run(k):
    net.embed(obs[k % 3], n_frames, out=out)
for i in range(3): run(i)
torch.cuda.synchronize()
def timeit(fn, reps=40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5): fn(i)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("eager pass ms:", timeit(run))
# graph: static input buffer
static = obs[0].clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    net.embed(static, n_frames, out=out)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    net.embed(static, n_frames, out=out)
def rung(k):
    static.copy_(obs[k % 3]); g.replay()
print("graph pass ms:", timeit(rung))
enc = net.encoder()
import numpy as np
ms = []
for r in range(6):
    net.transforms.run(obs[r % 3], n_frames, enc.slot0, enc.input_format, True)
    ms.append(sum(enc.forward_timed(out, net.out_size)))
print("sum of per-op ms:", np.mean(ms[1:]))
