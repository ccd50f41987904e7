"""Small embedding passes for compute-sanitizer (test infrastructure: uses the oracle's synthetic weights):
compute-sanitizer --tool memcheck python tests/sanitize_embed.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_parity as t
from oracle import restate
emb = np.load(os.path.join(ROOT, "tests", "golden", "embeddings.npz"))
for name, n, hw in (("moco_aug_uber_34", 3, 64), ("moco_aug", 2, 224)):
    net = t.make_net(name, emb["weight_seeds"])
    frames = restate.structured_frames(n, hw, hw, 3, 5)
    out = net.embed(torch.from_numpy(frames))
    torch.cuda.synchronize()
    print(name, tuple(out.shape), float(out.abs().mean()))
