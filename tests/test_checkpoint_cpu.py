"""Checkpoint / statistics schema of main_bc_2.py:250-258: the `.tar` written by bc.save_checkpoint has
the reference's keys and state_dict layouts; optimiser state interchanges with torch.optim.RMSprop both ways. CPU only
(constructors and state_dicts need no GPU; the arithmetic is covered by the -m gpu tests)."""
import argparse
import pickle

import pytest
import torch

from pvr_habitat_b200 import bc
from pvr_habitat_b200.embeddings import EmbeddingNet
from pvr_habitat_b200.models import PolicyNet
from pvr_habitat_b200.optim import FusedRMSprop


def _lr_lambda(max_epochs):
    return lambda epoch: 1 - min(epoch, max_epochs) / max_epochs  # main_bc_2.py:87-89


def test_reference_optimizer_state_loads_into_fused_rmsprop_and_back():
    torch.manual_seed(0)
    net = PolicyNet((32,), 3, batch_norm=True)
    ref_opt = torch.optim.RMSprop(net.parameters(), lr=1e-4, momentum=0, eps=0.01, alpha=0.99)  # main_bc_2.py:80-85
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    ref_opt.step()
    ref_opt.step()
    sd = ref_opt.state_dict()
    ours = FusedRMSprop(net.parameters(), lr=3e-4, eps=0.5, alpha=0.9, max_grad_norm=40.0)
    ours.load_state_dict(sd)
    g = ours.param_groups[0]
    assert (g["lr"], g["eps"], g["alpha"], g["momentum"]) == (1e-4, 0.01, 0.99, 0)
    for p in net.parameters():
        assert int(ours.state[p]["step"]) == 2
        assert torch.equal(ours.state[p]["square_avg"], ref_opt.state[p]["square_avg"])
    # and back: a state_dict written by the fused optimiser drives torch.optim.RMSprop
    for p in net.parameters():
        ours.state[p]["step"] = int(ours.state[p]["step"])  # what FusedRMSprop.step leaves behind
    back = torch.optim.RMSprop(net.parameters(), lr=1.0)
    back.load_state_dict(ours.state_dict())
    assert back.param_groups[0]["lr"] == 1e-4 and set(sd["param_groups"][0]) >= set(ours.state_dict()["param_groups"][0])
    before = [p.detach().clone() for p in net.parameters()]
    back.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
    with pytest.raises(NotImplementedError):
        FusedRMSprop(net.parameters(), momentum=0.9)


def test_checkpoint_roundtrip_has_reference_schema(tmp_path):
    torch.manual_seed(1)
    emb = EmbeddingNet("random", pretrained=False, disable_cuda=True)
    actor = PolicyNet((emb.out_size,), 3, batch_norm=True)
    opt = FusedRMSprop(actor.parameters(), lr=1e-4, eps=0.01, alpha=0.99, max_grad_norm=40.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, _lr_lambda(10))
    for p in actor.parameters():  # optimiser state as after one step
        opt.state[p]["step"] = 1
        opt.state[p]["square_avg"] = torch.rand_like(p)
    sched.step()
    flags = argparse.Namespace(env="HabitatImageNav-apartment_0", embedding_name="random", learning_rate=1e-4, run_id=3)
    stats = {"HabitatImageNav-apartment_0": {"frames": [0, 512], "training_loss": [-1.0, 1.09],
                                             "gradient_norm": [-1.0, 0.7]}}
    path = str(tmp_path / "run")
    bc.save_checkpoint(path, emb, actor, opt, sched, flags, stats)
    ck = torch.load(path + ".tar", map_location="cpu", weights_only=False)
    assert tuple(ck) == bc.CHECKPOINT_KEYS                       # key order of main_bc_2.py:252-258
    assert ck["flags"] == vars(flags)
    assert pickle.load(open(path + ".pickle", "rb")) == stats
    # reference module layouts: what torch's own modules of src/models.py:22-44 / src/embeddings.py:90-106 produce
    assert set(ck["actor_model_state_dict"]) >= {"fc.0.running_mean", "fc.1.weight", "fc.3.bias", "core.weight_hh_l1",
                                                 "policy.weight", "baseline.bias"}
    assert set(ck["embedding_model_state_dict"]) == {f"embedding.{i}.{k}" for i in (0, 2, 4, 6, 8)
                                                     for k in ("weight", "bias")}
    assert set(ck["actor_model_optimizer_state_dict"]["param_groups"][0]) >= {"lr", "momentum", "alpha", "eps",
                                                                               "centered", "weight_decay", "params"}
    # restore into fresh objects
    torch.manual_seed(2)
    emb2 = EmbeddingNet("random", pretrained=False, disable_cuda=True)
    actor2 = PolicyNet((emb.out_size,), 3, batch_norm=True)
    opt2 = FusedRMSprop(actor2.parameters(), lr=1.0)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, _lr_lambda(10))
    got = bc.load_checkpoint(path + ".tar", emb2, actor2, opt2, sched2)
    assert got == vars(flags)
    for (k, a), (_, b) in zip(actor.state_dict().items(), actor2.state_dict().items()):
        assert torch.equal(a, b), k
    for a, b in zip(emb.parameters(), emb2.parameters()):
        assert torch.equal(a, b)
    assert sched2.last_epoch == 1 and abs(opt2.param_groups[0]["lr"] - 1e-4 * 0.9) < 1e-12
    for p, q in zip(actor.parameters(), actor2.parameters()):
        assert torch.equal(opt.state[p]["square_avg"], opt2.state[q]["square_avg"])
    with pytest.raises(KeyError):
        torch.save({"model": {}}, path + ".bad")
        bc.load_checkpoint(path + ".bad")


def test_lstm_wavefront_chunk_count(monkeypatch):
    """PolicyNet._lstm_chunks: the environment override, halving until the chunks divide T with at least two steps
    each, and the eager default (1: the two-stream wavefront is only the default inside a captured step)."""
    net = PolicyNet((16,), 3)
    monkeypatch.delenv("PVR_LSTM_CHUNKS", raising=False)
    assert [net._lstm_chunks(t) for t in (1, 7, 64, 100)] == [1, 1, 1, 1]      # not capturing on CPU
    monkeypatch.setenv("PVR_LSTM_CHUNKS", "8")
    assert [net._lstm_chunks(t) for t in (1, 2, 7, 8, 16, 64, 100, 12)] == [1, 1, 1, 4, 8, 8, 4, 4]
    monkeypatch.setenv("PVR_LSTM_CHUNKS", "16")
    assert net._lstm_chunks(64) == 16 and net._lstm_chunks(16) == 8
    monkeypatch.setenv("PVR_LSTM_CHUNKS", "1")
    assert net._lstm_chunks(64) == 1
