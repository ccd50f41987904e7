"""CPU: the flags of the BC entry scripts against the reference's parser (tests/golden/arguments.json, generated from
src/arguments.py by oracle/make_golden_bc.py) and the host logic of pvr_habitat_b200.bc_run (file names, statistics
schema, resume / already-complete behaviour, essential-save NaN filling) with the CUDA trainer replaced by a stub."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from pvr_habitat_b200 import bc_run
from pvr_habitat_b200.arguments import make_parser


def test_parser_matches_reference(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "arguments.json")))
    mine = {a.dest: a for a in make_parser()._actions if a.dest != "help"}
    assert set(mine) == set(spec)
    for dest, ref in spec.items():
        a = mine[dest]
        assert list(a.option_strings) == ref["flags"], dest
        assert a.default == ref["default"], dest
        assert getattr(a.type, "__name__", None) == ref["type"], dest
        if a.nargs == 0:
            assert a.const == ref["const"], dest


class _StubTrainer:
    """Stands in for BCTrainer: deterministic fake losses, records its construction."""
    made = []

    def __init__(self, actor_model, obs, action, done, batch_size, unroll_length, max_frames, **kw):
        self.kw, self.T, self.B = kw, unroll_length, batch_size
        self.n, self.frames, self.k = len(action), 0, 0
        self.max_epochs = max_frames // (batch_size * unroll_length) + 1
        self.optimizer = torch.optim.RMSprop(actor_model.parameters(), lr=kw.get("learning_rate", 1e-4))
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lambda e: 1 - e / self.max_epochs)
        _StubTrainer.made.append(self)

    def step(self):
        self.k += 1
        self.frames += self.T * self.B
        return torch.tensor(1.0 / self.k)

    def gradient_norm(self):
        return torch.tensor(2.0 * self.k)


class _StubPolicy(torch.nn.Module):
    def __init__(self, obs_shape, num_actions, batch_norm=False):
        super().__init__()
        self.fc = torch.nn.Linear(int(np.prod(obs_shape)), num_actions)
        self.obs_shape, self.num_actions, self.batch_norm = tuple(obs_shape), num_actions, batch_norm


class _StubEmbedding(torch.nn.Module):
    out_size = 5

    def __init__(self, name, **kw):
        super().__init__()
        self.name = name
        self.embedding = torch.nn.Linear(2, 2)

    def embed(self, obs, n_frames):
        return obs.float().mean((1, 2)).reshape(len(obs), n_frames, 3)[..., :1].repeat(1, 1, 5).reshape(len(obs), -1)


@pytest.fixture
def patched(monkeypatch):
    _StubTrainer.made.clear()
    monkeypatch.setattr(bc_run, "BCTrainer", _StubTrainer)
    monkeypatch.setattr(bc_run, "PolicyNet", _StubPolicy)
    monkeypatch.setattr(bc_run, "PolicyNetWithConv", _StubPolicy)
    monkeypatch.setattr(bc_run, "EmbeddingNet", _StubEmbedding)
    monkeypatch.setattr(bc_run, "_device", lambda flags: torch.device("cpu"))


def _flags(tmp_path, *extra):
    return make_parser().parse_args([
        "--env", "sceneA,sceneB", "--to_env", "sceneC", "--embedding_name", "emb", "--data_path", str(tmp_path),
        "--save_path", str(tmp_path / "out"), "--batch_size", "2", "--unroll_length", "4", "--max_frames", "80",
        "--eval_frequency", "2", "--run_id", "7", *extra])


def _write_embedded(tmp_path, n=30, d=6):
    for i, env in enumerate(("sceneA", "sceneB")):
        rng = np.random.default_rng(i)
        with open(tmp_path / f"{env}_emb.pickle", "wb") as f:
            pickle.dump(dict(obs=rng.standard_normal((n + i, d)).astype(np.float32), action=rng.integers(0, 3, n + i),
                             reward=np.zeros(n + i, np.float32), done=rng.random(n + i) < 0.1,
                             true_state=np.zeros((n + i, 12))), f)


def test_bc2_files_stats_and_resume(tmp_path, patched):
    _write_embedded(tmp_path)
    stats = bc_run.run_bc(_flags(tmp_path), "bc2")
    tr = _StubTrainer.made[-1]
    assert tr.n == 61 and tr.kw["max_grad_norm"] == 40.0 and tr.kw["alpha"] == 0.99  # both scenes concatenated
    base = tmp_path / "out" / "sceneA,sceneB_ememb_s7_sceneC"  # main_bc_2.py:43-47
    assert os.path.isfile(str(base) + ".pickle") and os.path.isfile(str(base) + ".tar")
    st = stats["sceneC"]
    assert set(st) == {"episode_return", "episode_success", "frames", "training_loss", "gradient_norm"}
    # 10 steps of 8 frames, evaluation every 2nd step: entries at frames 8, 24, ... (the `frames` of the step itself)
    assert st["frames"] == [0, 8, 24, 40, 56, 72] and np.isnan(st["training_loss"][0])
    assert np.allclose(st["training_loss"][1:], [1 / 2, 1 / 4, 1 / 6, 1 / 8, 1 / 10])
    assert all(np.isnan(v) for v in st["episode_success"])  # no simulator injected
    ck = torch.load(str(base) + ".tar", weights_only=False)
    assert list(ck) == ["embedding_model_state_dict", "actor_model_state_dict", "actor_model_optimizer_state_dict",
                        "scheduler_state_dict", "flags"] and ck["flags"]["unroll_length"] == 4
    # a run whose last saved `frames` reaches max_frames returns immediately (main_bc_2.py:52-55) ...
    n_made = len(_StubTrainer.made)
    bc_run.run_bc(_flags(tmp_path, "--max_frames", "72"), "bc2")
    assert len(_StubTrainer.made) == n_made
    # ... a longer max_frames resumes from the last saved `frames` (the step at that value runs again, :162,183)
    stats2 = bc_run.run_bc(_flags(tmp_path, "--max_frames", "104"), "bc2")
    assert stats2["sceneC"]["frames"] == [0, 8, 24, 40, 56, 72, 72, 88]
    assert _StubTrainer.made[-1].frames == 72 + 4 * 8


def test_injected_environment_and_rollouts(tmp_path, patched):
    _write_embedded(tmp_path)
    calls = []

    def make_environment(flags, embedding_model=None):
        calls.append(flags.env)
        return bc_run._DatasetEnv((6,), 3)

    def test(model, env, stat_keys, n_episodes):
        return {k: [1.0, 0.0] for k in stat_keys}

    st = bc_run.run_bc(_flags(tmp_path, "--essential_save_only"), "bc2", make_environment, test)["sceneC"]
    assert calls == ["sceneC"]  # flags.env = to_env before the environment is made (main_bc_2.py:74-75)
    assert st["episode_success"][0] == 0.5 and len(st["episode_success"]) == len(st["frames"])


def test_bc1_and_finetune_front_ends(tmp_path, patched):
    rng = np.random.default_rng(0)
    traj = lambda L: rng.integers(0, 255, (L, 8, 8, 6), dtype=np.uint8)  # noqa: E731
    for env in ("sceneA", "sceneB"):
        with open(tmp_path / f"{env}.pickle", "wb") as f:
            pickle.dump(dict(obs=[traj(9), traj(11)], action=[np.zeros(9, int), np.ones(11, int)],
                             reward=[np.zeros(9), np.zeros(11)], done=[np.zeros(9, bool), np.zeros(11, bool)],
                             true_state=[np.zeros((9, 12)), np.zeros((11, 12))]), f)
    bc_run.run_bc(_flags(tmp_path), "bc1")
    assert os.path.isfile(tmp_path / "out" / "sceneA,sceneB_ememb_s7_sceneC.tar")
    bc_run.run_bc(_flags(tmp_path), "finetune")
    base = tmp_path / "out" / "sceneA,sceneB_emrandom_finetuned_s7_sceneC"  # main_bc_finetune.py:42-46
    ck = torch.load(str(base) + ".tar", weights_only=False)
    assert "embedding_model_state_dict" not in ck  # main_bc_finetune.py:232-238
    assert _StubTrainer.made[-1].n == 40 and _StubTrainer.made[-2].n == 40
