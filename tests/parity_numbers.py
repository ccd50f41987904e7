"""Print rel-L2 / cosine of the CUDA embeddings against the committed reference goldens (tests/golden/embeddings.npz)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as t  # noqa: E402

emb = np.load(os.path.join(ROOT, "tests", "golden", "embeddings.npz"))
for name, mode in [(n, m) for m in ("bf16", "fp32") for n in t.VARIANTS]:
    net = t.make_net(name, emb["weight_seeds"]).set_precision(mode)
    for tag in ("64", "224"):
        key = f"emb{tag}_{name}"
        if key not in emb.files:
            continue
        got = net(torch.from_numpy(emb["frames" + tag])).astype(np.float64)
        ref = emb[key].astype(np.float64)
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        cos = min(float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(got, ref))
        print(f"{name:20s} {mode} {tag:>3s}x{tag:<3s} rel-L2 {rel:.2e}  min cosine {cos:.9f}")
