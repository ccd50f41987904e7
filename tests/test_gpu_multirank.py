"""-m gpu: data-parallel BC on the CUDA path, 2 ranks launched with torch.distributed.run, against the single-process
run of the same GLOBAL batch (main_bc_2.py:186-227 / main_bc_finetune.py:167-208 at the global batch size; the
reference has no distributed code). With >= 2 GPUs the ranks run on separate devices over NCCL; on a
one-GPU box both ranks share the device and the collectives go through gloo — the same product code (BCTrainer,
PolicyNet(WithConv) BatchNorm-sum and gradient all-reduces, parameter broadcast) either way.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import multirank_worker as mw

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ranks(case, out, world=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py"), case, str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.load(open(out))


# (finetune: the conv trunk's weight gradients are split-K sums whose split depends on the number of rows per rank)
@pytest.mark.parametrize("case,tol", [("policy", 1e-4), ("finetune", 5e-4)])
def test_dp2_trace_equals_single_process(tmp_path, case, tol):
    dp = _run_ranks(case, tmp_path / "dp.json")
    single = mw.train(case, torch.device("cuda"), keep_state=True, use_graph=False)
    assert dp["world"] == 2 and single["world"] == 1
    l1, l2 = np.array(single["loss"]), np.array(dp["loss"])
    n1, n2 = np.array(single["grad_norm"]), np.array(dp["grad_norm"])
    print(case, dp["backend"], "loss", l1[:3], l2[:3], "max rel diff", np.abs(l1 - l2).max() / l1.max(),
          "grad-norm max rel diff", (np.abs(n1 - n2) / n1).max())
    # bf16 GEMMs over a half batch vs the full batch differ only in fp32 summation order
    assert np.allclose(l1, l2, rtol=tol, atol=tol), (l1, l2)
    assert np.allclose(n1, n2, rtol=50 * tol), (n1, n2)
    # replicas stay identical (same all-reduced gradients, parameters broadcast from rank 0 at the start) ...
    a, b = dp["replica_param_sums"]
    assert all(a[k] == b[k] for k in a), "replicas diverged"
    # ... and close to the single-process parameters after the same number of steps. (RMSprop's first steps are
    # sign-like — |update| ~ 10 lr whatever the gradient's size — so the 1e-5 differences of the summation order move
    # individual small-gradient coordinates by a full step: the bound is on the tensors' relative L2 distance.)
    dp_state = torch.load(str(tmp_path / "dp.json") + ".state.pt")
    # (weight matrices only: zero-initialised biases consist of nothing but those few sign-like steps)
    worst = max(float((dp_state[k] - v).norm() / (v.norm() + 1e-12)) for k, v in single["state"].items()
                if v.dim() >= 2 and v.numel() >= 1024)
    print("largest per-tensor relative L2 distance of the weight matrices after training:", worst)
    assert worst < 5e-2
