"""ORACLE (build container only) — loss-curve fixtures of the UNMODIFIED reference entry scripts.

    python oracle/make_golden_bc.py config5      # main_bc_2.run at BASELINE configs[4]'s shape -> bc_config5_curve.npz
    python oracle/make_golden_bc.py finetune     # main_bc_finetune.run (configs[3])            -> bc_finetune_curve.npz
    python oracle/make_golden_bc.py bc1          # main_bc_1.run ('random' PVR -> BC in RAM)     -> bc1_curve.npz

Each run executes the reference's own `run(flags)` (fake `src.env_utils` / `src.test_model` modules: the simulator is
out of scope) on a synthetic dataset that the tests regenerate from the stored seed
(`oracle.restate_policy.synthetic_bc_data` / `oracle.restate.structured_frames`), with `--eval_frequency 1` so that
the statistics pickle holds the loss and the pre-clip gradient norm of EVERY step. Only the traces are stored.
"""
import os
import pickle
import sys
import tempfile
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, restate  # noqa: E402
from oracle import restate_policy as rp  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def fake_modules(obs_shape):
    """src.env_utils / src.test_model stand-ins: the observation shape and 3 actions (src/gym_wrappers.py:173) are all
    the BC scripts read from the environment; `test` (simulator rollouts) returns zeros."""
    env_utils = types.ModuleType("src.env_utils")

    class _Space:
        def __init__(self, shape=None, n=None):
            self.shape, self.n = shape, n

    class _Env:
        def __init__(self):
            self.gym_env = types.SimpleNamespace(observation_space=_Space(shape=tuple(obs_shape)),
                                                 action_space=_Space(n=3))

        def close(self):
            pass

    env_utils.make_environment = lambda flags, embedding_model=None, actor_id=1: _Env()
    test_model = types.ModuleType("src.test_model")
    test_model.test = lambda model, env, stat_keys, n_episodes: {k: [0.0] for k in stat_keys}
    sys.modules["src.env_utils"], sys.modules["src.test_model"] = env_utils, test_model
    refshim.install_stubs()


def read_stats(path, to_env):
    st = pickle.load(open(path, "rb"))[to_env]
    return dict(loss=np.array(st["training_loss"][1:], dtype=np.float64),
                grad_norm=np.array(st["gradient_norm"][1:], dtype=np.float64), frames=np.array(st["frames"][1:]))


def config5(steps=120, n=16384, D=2048, T=64, B=128, data_seed=21, run_id=3):
    """BASELINE configs[4]: main_bc_2 on pre-embedded 2048-d observations, global batch T*B = 8192, BatchNorm on."""
    obs, action, done, reward = rp.synthetic_bc_data(n, D, 3, data_seed)
    fake_modules((D,))
    import src.embeddings as E
    E.EmbeddingNet = lambda *a, **k: torch.nn.Identity()
    import main_bc_2
    main_bc_2.EmbeddingNet = E.EmbeddingNet
    from src.arguments import parser
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "fakeenv_fakeemb.pickle"), "wb") as f:
            pickle.dump(dict(obs=obs, action=action, reward=reward, done=done, true_state=np.zeros((n, 12))), f)
        flags = parser.parse_args([
            "--env", "fakeenv", "--to_env", "fakeenv", "--embedding_name", "fakeemb", "--data_path", d,
            "--save_path", os.path.join(d, "out"), "--batch_size", str(B), "--unroll_length", str(T),
            "--max_frames", str(steps * T * B), "--eval_frequency", "1", "--batch_norm", "--disable_cuda",
            "--run_id", str(run_id), "--n_episodes_test", "1"])
        main_bc_2.run(flags)
        st = read_stats(os.path.join(d, "out", f"fakeenv_emfakeemb_s{run_id}_fakeenv.pickle"), "fakeenv")
        ck = torch.load(os.path.join(d, "out", f"fakeenv_emfakeemb_s{run_id}_fakeenv.tar"), weights_only=False)
    sd = ck["actor_model_state_dict"]
    print("config5:", steps, "steps in", round(time.time() - t0), "s; loss", st["loss"][:3], "...", st["loss"][-3:])
    np.savez_compressed(
        os.path.join(GOLDEN, "bc_config5_curve.npz"), n=n, D=D, T=T, B=B, steps=steps, data_seed=data_seed,
        run_id=run_id, max_frames=steps * T * B, **st,
        final_param_sums=np.array([float(v.double().sum()) for v in sd.values()]),
        final_param_names=np.array(list(sd.keys())),
        final_policy_weight=sd["policy.weight"].numpy(), final_fc_bias=sd["fc.1.bias"].numpy())


def finetune(steps=40, n_traj=24, traj_len=80, T=20, B=16, data_seed=33, run_id=4):
    """BASELINE configs[3]: main_bc_finetune on raw 64x64 2-frame observations (list-over-trajectories pickle,
    behavioral_cloning/save_opt_trajectories.py:94-106)."""
    data, action = rp.synthetic_frame_trajectories(n_traj, traj_len, data_seed)
    fake_modules((64, 64, 6))
    import main_bc_finetune
    from src.arguments import parser
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "fakeenv.pickle"), "wb") as f:
            pickle.dump(data, f)
        flags = parser.parse_args([
            "--env", "fakeenv", "--to_env", "fakeenv", "--data_path", d, "--save_path", os.path.join(d, "out"),
            "--batch_size", str(B), "--unroll_length", str(T), "--max_frames", str(steps * T * B),
            "--eval_frequency", "1", "--batch_norm", "--disable_cuda", "--run_id", str(run_id),
            "--n_episodes_test", "1"])
        main_bc_finetune.run(flags)
        st = read_stats(os.path.join(d, "out", f"fakeenv_emrandom_finetuned_s{run_id}_fakeenv.pickle"), "fakeenv")
    print("finetune:", steps, "steps in", round(time.time() - t0), "s; loss", st["loss"][:3], "...", st["loss"][-3:])
    np.savez_compressed(os.path.join(GOLDEN, "bc_finetune_curve.npz"), n_traj=n_traj, traj_len=traj_len, T=T, B=B,
                        steps=steps, data_seed=data_seed, run_id=run_id, max_frames=steps * T * B,
                        action=action, **st)


def bc1(steps=30, n_traj=16, traj_len=64, T=16, B=8, data_seed=35, run_id=6):
    """main_bc_1.run with the 'random' PVR (weights depend on the seed: src/embeddings.py:90-106): raw frames ->
    EmbeddingNet in mini-batches -> BC on the embeddings kept in RAM (main_bc_1.py:117-138, 193-234)."""
    data, action = rp.synthetic_frame_trajectories(n_traj, traj_len, data_seed)
    E = refshim.reference_embeddings()
    torch.manual_seed(run_id)
    probe = E.EmbeddingNet("random", pretrained=True, train=False, disable_cuda=True)
    fake_modules((probe.out_size * 2,))
    import main_bc_1
    from src.arguments import parser
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "fakeenv.pickle"), "wb") as f:
            pickle.dump(data, f)
        flags = parser.parse_args([
            "--env", "fakeenv", "--to_env", "fakeenv", "--embedding_name", "random", "--data_path", d,
            "--save_path", os.path.join(d, "out"), "--batch_size", str(B), "--unroll_length", str(T),
            "--max_frames", str(steps * T * B), "--eval_frequency", "1", "--batch_norm", "--disable_cuda",
            "--run_id", str(run_id), "--n_episodes_test", "1"])
        main_bc_1.run(flags)
        st = read_stats(os.path.join(d, "out", f"fakeenv_emrandom_s{run_id}_fakeenv.pickle"), "fakeenv")
        ck = torch.load(os.path.join(d, "out", f"fakeenv_emrandom_s{run_id}_fakeenv.tar"), weights_only=False)
    emb_keys = sorted(ck["embedding_model_state_dict"].keys())
    print("bc1:", steps, "steps in", round(time.time() - t0), "s; loss", st["loss"][:3], "...", st["loss"][-3:])
    np.savez_compressed(os.path.join(GOLDEN, "bc1_curve.npz"), n_traj=n_traj, traj_len=traj_len, T=T, B=B, steps=steps,
                        data_seed=data_seed, run_id=run_id, max_frames=steps * T * B, action=action,
                        embedding_keys=np.array(emb_keys),
                        embedding_conv0_sum=float(ck["embedding_model_state_dict"][emb_keys[0]].double().sum()), **st)


def args():
    """Names, defaults and types of the reference's shared parser (src/arguments.py:3-68)."""
    import json
    refshim.install_stubs()
    from src.arguments import parser
    spec = {a.dest: dict(default=a.default, type=getattr(a.type, "__name__", None), flags=list(a.option_strings),
                         const=a.const if a.nargs == 0 else None)
            for a in parser._actions if a.dest != "help"}
    with open(os.path.join(GOLDEN, "arguments.json"), "w") as f:
        json.dump(spec, f, indent=1, sort_keys=True)
    print("args:", len(spec), "flags")


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", "8")))
    os.makedirs(GOLDEN, exist_ok=True)
    for which in sys.argv[1:] or ["config5", "finetune", "bc1"]:
        {"config5": config5, "finetune": finetune, "bc1": bc1, "args": args}[which]()
