"""Small embedding passes for compute-sanitizer: compute-sanitizer --tool memcheck python tools/sanitize_embed.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_parity as t
from oracle import restate
emb = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "embeddings.npz"))
for name, n, hw in (("moco_aug_uber_34", 3, 64), ("moco_aug", 2, 224)):
    net = t.make_net(name, emb["weight_seeds"])
    frames = restate.structured_frames(n, hw, hw, 3, 5)
    out = net.embed(torch.from_numpy(frames))
    torch.cuda.synchronize()
    print(name, tuple(out.shape), float(out.abs().mean()))
