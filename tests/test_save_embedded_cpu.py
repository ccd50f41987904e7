"""Host logic of pvr_habitat_b200/save_embedded_obs.py (drop-in for behavioral_cloning/save_embedded_obs.py) against the
artefacts the UNMODIFIED reference wrote for the same synthetic trajectories (tests/golden/save_embedded.npz, made by
oracle/make_golden.py): file names, pickle keys / order / dtypes, sample order, frame regrouping, the .tar schema.
The encoder is replaced by a CPU stand-in built on the oracle (the CUDA `EmbeddingNet.embed` has its own -m gpu parity
tests), so this runs without a GPU; a two-rank gloo run checks the sharded variant."""
import argparse
import os
import pickle

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import restate
from pvr_habitat_b200 import save_embedded_obs as S


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "save_embedded.npz"))


class OracleEncoder:
    """Stands in for EmbeddingNet('random'): same surface as far as run() uses it."""
    calls = []

    def __init__(self, weights):
        self.sd = {k: torch.from_numpy(v) for k, v in weights.items()}
        self.out_size = 1568

    def __call__(self, embedding_name, in_channels=3, pretrained=True, train=False, disable_cuda=False):
        assert (embedding_name, in_channels, train) == ("random", 3, False)
        return self

    def state_dict(self):
        return {"embedding." + k: v for k, v in self.sd.items()}

    def embed(self, obs, n_frames):
        OracleEncoder.calls.append(tuple(obs.shape))
        frames, n = restate.split_frames(obs.numpy())
        assert n == n_frames
        return torch.from_numpy(restate.regroup_frames(restate.small_conv_embedding(self.sd, frames), n_frames))


def write_inputs(gold, d):
    env = str(gold["env"])
    cuts = np.cumsum(gold["lengths"])[:-1]
    traj = {k: np.split(gold["in_" + k], cuts) for k in ("obs", "action", "reward", "done", "true_state")}
    with open(os.path.join(d, env + ".pickle"), "wb") as fh:
        pickle.dump(traj, fh, protocol=pickle.HIGHEST_PROTOCOL)
    return env, traj


def flags_for(gold, d, source, **kw):
    ns = argparse.Namespace(data_path=str(d), env=str(gold["env"]), embedding_name="random", run_id=int(gold["run_id"]),
                            batch_size=4, disable_cuda=True, pretrained_embedding=True, train_embedding=False,
                            source=source, n_trajectories=-1)
    ns.__dict__.update(kw)
    return ns


def encoder(gold):
    return OracleEncoder({k[len("w_embedding."):]: gold[k] for k in gold.files if k.startswith("w_embedding.")})


def check_pickle(gold, source, data, d):
    assert list(data.keys()) == list(gold[f"{source}_keys"])
    for k in ("action", "reward", "done", "true_state"):
        ref = gold[f"{source}_{k}"]
        assert data[k].dtype == ref.dtype and np.array_equal(data[k], ref), k
    ref = gold[f"{source}_obs"]
    assert data["obs"].shape == ref.shape and data["obs"].dtype == np.float32
    np.testing.assert_allclose(data["obs"], ref, rtol=1e-4, atol=1e-5 * float(np.abs(ref).max()))
    if source == "png":
        assert [os.path.relpath(p, d) for p in data["png"]] == list(gold["png_png"])


def test_pickle_source_reproduces_reference_artefacts(gold, tmp_path):
    env, _ = write_inputs(gold, tmp_path)
    OracleEncoder.calls.clear()
    S.run(flags_for(gold, tmp_path, "pickle"), embedding_factory=encoder(gold))
    files = sorted(f for f in os.listdir(tmp_path) if os.path.isfile(tmp_path / f))
    # the generator listed the directory after moving the result away: <env>.pickle, random_<run_id>.tar (+ the result)
    assert files == sorted(list(gold["files"]) + [env + "_random.pickle"])
    with open(tmp_path / (env + "_random.pickle"), "rb") as fh:
        check_pickle(gold, "pickle", pickle.load(fh), tmp_path)
    ck = torch.load(tmp_path / f"random_{int(gold['run_id'])}.tar", map_location="cpu")
    assert list(ck.keys()) == list(gold["tar_keys"])
    assert set(ck["embedding_model_state_dict"]) == {k[2:] for k in gold.files if k.startswith("w_embedding.")}
    assert OracleEncoder.calls == [(12, 64, 64, 6)]           # one fused pass instead of three mini-batches of 4
    # an existing result is not recomputed (save_embedded_obs.py:99-100)
    OracleEncoder.calls.clear()
    S.run(flags_for(gold, tmp_path, "pickle"), embedding_factory=encoder(gold))
    assert OracleEncoder.calls == []


def test_png_source_reproduces_reference_artefacts(gold, tmp_path):
    cv2 = pytest.importorskip("cv2")
    env, traj = write_inputs(gold, tmp_path)
    os.makedirs(tmp_path / env)
    for t, frames in enumerate(traj["obs"]):                   # the layout save_opt_trajectories_png.py:43-58 writes
        for s in range(len(frames)):
            cv2.imwrite(str(tmp_path / env / f"{t}_{s}.png"), frames[s][:, :, :3])
        cv2.imwrite(str(tmp_path / env / f"{t}_goal.png"), frames[-1][:, :, 3:])
        with open(tmp_path / env / f"{t}.pickle", "wb") as fh:
            pickle.dump({k: traj[k][t] for k in ("action", "reward", "done", "true_state")}, fh)
    OracleEncoder.calls.clear()
    S.run(flags_for(gold, tmp_path, "png"), embedding_factory=encoder(gold))
    with open(tmp_path / (env + "_random.pickle"), "rb") as fh:
        check_pickle(gold, "png", pickle.load(fh), tmp_path)
    # goal + all frames of a trajectory in one encoder call (the reference: one call per image)
    assert OracleEncoder.calls == [(n + 1, 64, 64, 3) for n in gold["lengths"]]


def test_read_pickle_merges_trajectories(gold, tmp_path):
    env, traj = write_inputs(gold, tmp_path)
    data = S.read_habitat_data_from_pickle(str(tmp_path / env))
    assert np.array_equal(data["obs"], gold["in_obs"]) and np.array_equal(data["done"], gold["in_done"])
    two = S.read_habitat_data_from_pickle(str(tmp_path / env), n_trajectories=2)
    assert len(two["reward"]) == int(gold["lengths"][:2].sum())


def _rank_main(rank, world, port, d, gold_path):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = np.load(gold_path)
    S.run(flags_for(g, d, "pickle"), embedding_factory=encoder(g))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_process(gold, golden_dir, tmp_path):
    """Samples cut into contiguous blocks per rank (7 + 5... here 6 + 6 of 12), gathered on rank 0, one set of files."""
    env, _ = write_inputs(gold, tmp_path)
    port = 29650 + os.getpid() % 200
    mp.spawn(_rank_main, args=(2, port, str(tmp_path), os.path.join(golden_dir, "save_embedded.npz")), nprocs=2, join=True)
    with open(tmp_path / (env + "_random.pickle"), "rb") as fh:
        check_pickle(gold, "pickle", pickle.load(fh), tmp_path)
    assert sorted(f for f in os.listdir(tmp_path) if os.path.isfile(tmp_path / f)) == \
        sorted(list(gold["files"]) + [env + "_random.pickle"])
