"""The behavioural-cloning entry scripts of the reference — main_bc_2.run (pre-embedded observations),
main_bc_1.run (raw frames -> frozen PVR -> BC on embeddings kept in memory) and main_bc_finetune.run (raw frames,
conv trunk trained end to end) — as one loop with three data front-ends. `run(flags)` takes the reference's flags
(pvr_habitat_b200.arguments), reads the reference's pickles, writes the reference's `.pickle` statistics / `.tar`
checkpoint pair under the same names, resumes from them the same way, and trains on the CUDA path (BCTrainer).

What is NOT here: the Habitat simulator. The reference gets the observation shape and the number of actions from
`make_environment(flags, embedding_model)` and evaluates with rollouts `test(test_model, env, ...)`
(main_bc_2.py:75-77, 176, 236). Both are injected: `run(flags, make_environment=..., test=...)`. Without them the
observation shape comes from the dataset, the action count is Habitat's 3 (src/gym_wrappers.py:173) and the
`episode_return` / `episode_success` statistics are NaN (the layout the reference itself uses for skipped evaluations,
main_bc_2.py:240-242).

Differences by design: the embedding pass of main_bc_1 goes through `EmbeddingNet.embed` in large passes and stays on
the device — the (N, O*n) table never visits the host (second half); the per-step host gather + H2D of
T*B*D floats is a device gather driven by the same `sample_with_minimum_distance` draws; gradient-norm / clip / RMSprop
are the fused kernels (pvr_habitat_b200.optim). Under torchrun every rank draws the same starts and trains its slice
of the global batch (parallel.py); rank 0 writes the files.
"""
import os
import pickle
import random

import numpy as np
import torch

from . import parallel
from .bc import BCTrainer, save_checkpoint
from .embeddings import EmbeddingNet
from .models import PolicyNet, PolicyNetWithConv
from .utils_bc import is_essential_save

STAT_KEYS = ['episode_return', 'episode_success']
NUM_ACTIONS = 3  # src/gym_wrappers.py:173 (Habitat: forward / left / right)


class _DatasetEnv:
    """Stand-in for the simulator environment when none is injected: only what the BC scripts read from it."""

    class _Space:
        def __init__(self, shape=None, n=None):
            self.shape, self.n = shape, n

    def __init__(self, obs_shape, num_actions=NUM_ACTIONS):
        self.gym_env = type("GymEnv", (), {})()
        self.gym_env.observation_space = self._Space(shape=tuple(obs_shape))
        self.gym_env.action_space = self._Space(n=num_actions)

    def close(self):
        pass


def _seed(flags):
    torch.manual_seed(flags.run_id)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(flags.run_id)
    np.random.seed(flags.run_id)
    random.seed(flags.run_id)


def _device(flags):
    if flags.disable_cuda or not torch.cuda.is_available():
        raise RuntimeError("pvr_habitat_b200: the BC scripts need CUDA (there is no CPU path); drop --disable_cuda")
    rank, world, local = parallel.env_world()
    return torch.device('cuda', local if world > 1 else torch.cuda.current_device())


def init_distributed_from_env():
    """Under torchrun (WORLD_SIZE > 1) join the NCCL process group, one rank per GPU; no-op for a single process."""
    rank, world, local = parallel.env_world()
    if world > 1 and not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world


def _is_writer():
    return not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
        torch.distributed.get_rank() == 0


def _merge_trajectories(data, keys=('obs', 'action', 'reward', 'done', 'true_state')):
    """src/utils_bc.py:33-49: lists over trajectories -> arrays over steps."""
    n_traj = len(data['reward'])
    for k in keys:
        if k in data:
            data[k] = np.concatenate(data[k])
    print('  ', '%d trajectories for a total of %d samples' % (n_traj, len(data['reward'])))
    print('  ', 'avg. return is', data['reward'].sum() / max(n_traj, 1))
    return data


# ------------------------------------------------------------------------------------------------- data front-ends
def load_embedded(flags, envs, embedding_model, device):
    """main_bc_2.py:111-143: <data_path>/<env>_<embedding>.pickle, flat over steps."""
    cols = dict(obs=[], action=[], reward=[], done=[])
    for env_id in envs:
        name = env_id + ('_resnet50' if flags.embedding_name == 'true_state' else '_' + flags.embedding_name)
        with open(os.path.join(flags.data_path, name + '.pickle'), 'rb') as fh:
            data = pickle.load(fh)
        n = flags.batch_size * flags.unroll_length if flags.debug else data['obs'].shape[0]
        cols['obs'].append(np.asarray(data['true_state' if flags.embedding_name == 'true_state' else 'obs'][:n]))
        for k in ('action', 'reward', 'done'):
            cols[k].append(np.asarray(data[k][:n]))
    return {k: np.concatenate(v) for k, v in cols.items()}


def load_and_embed(flags, envs, embedding_model, device, pass_size=2048):
    """main_bc_1.py:112-150: raw trajectories, one scene at a time, through the frozen encoder. The reference embeds
    mini-batches of `batch_size` on the host loop; here passes of `pass_size` observations go through
    `EmbeddingNet.embed` and the rows are written straight into the device-resident (N, O*n) table."""
    tables, cols = [], dict(action=[], reward=[], done=[])
    for env_id in envs:
        print('loading %s ...' % os.path.join(flags.data_path, env_id + '.pickle'))
        with open(os.path.join(flags.data_path, env_id + '.pickle'), 'rb') as fh:
            data = _merge_trajectories(pickle.load(fh))
        n = flags.batch_size * flags.unroll_length if flags.debug else data['obs'].shape[0]
        obs = data['obs'][:n]
        if obs.shape[-1] == 1:  # grayscale (main_bc_1.py:132-133)
            obs = np.repeat(obs, 3, -1)
        n_frames = max(obs.shape[3] // 3, 1)
        print('  ', 'passing observations through embedding model')
        table = torch.empty(len(obs), embedding_model.out_size * n_frames, dtype=torch.float32, device=device)
        for i in range(0, len(obs), pass_size):
            o = torch.from_numpy(np.ascontiguousarray(obs[i:i + pass_size]))
            table[i:i + pass_size] = embedding_model.embed(o, n_frames)
        tables.append(table)
        for k in cols:
            cols[k].append(np.asarray(data[k][:n]))
    out = {k: np.concatenate(v) for k, v in cols.items()}
    out['obs'] = torch.cat(tables) if len(tables) > 1 else tables[0]
    return out


def load_raw(flags, envs, embedding_model, device):
    """main_bc_finetune.py:104-124: raw uint8 frames; --debug truncates to that many TRAJECTORIES (:109-112)."""
    cols = dict(obs=[], action=[], reward=[], done=[])
    for env_id in envs:
        with open(os.path.join(flags.data_path, env_id + '.pickle'), 'rb') as fh:
            data = pickle.load(fh)
        n = flags.batch_size * flags.unroll_length if flags.debug else len(data['obs'])
        for k in cols:
            cols[k].append(np.concatenate(data[k][:n]))
    return {k: np.concatenate(v) for k, v in cols.items()}


# ------------------------------------------------------------------------------------------------- the loop
def run_bc(flags, mode, make_environment=None, test=None, trainer_kwargs=None):
    """mode: 'bc2' (main_bc_2), 'bc1' (main_bc_1), 'finetune' (main_bc_finetune). Returns the statistics dict."""
    _seed(flags)
    if flags.debug:
        flags.n_episodes_test = int(np.minimum(2, flags.n_episodes_test))
    from_env, to_env = flags.env, flags.to_env
    os.makedirs(flags.save_path, exist_ok=True)
    tag = 'random_finetuned' if mode == 'finetune' else flags.embedding_name
    save_path = os.path.join(flags.save_path, from_env + '_em' + tag + '_s' + str(flags.run_id) + '_' + to_env)

    resume = False
    if os.path.isfile(save_path + '.pickle'):  # main_bc_2.py:49-56
        with open(save_path + '.pickle', 'rb') as fh:
            stats = pickle.load(fh)
        if stats[to_env]['frames'][-1] >= flags.max_frames:
            print('   WARNING! This run was already completed. Stopping now.')
            return stats
        resume = True

    flags.device = _device(flags)
    embedding_model = None
    if mode != 'finetune':  # same construction order as the reference: the encoder draws from the torch RNG first
        embedding_model = EmbeddingNet(flags.embedding_name, in_channels=3, pretrained=True, train=False,
                                       disable_cuda=flags.disable_cuda)
    loader = dict(bc2=load_embedded, bc1=load_and_embed, finetune=load_raw)[mode]

    print('=== Loading trajectories ===')
    # (the reference builds the models first; nothing below consumes the seeded RNG streams, so reading the data before
    # the policy is constructed does not change a single draw)
    data = loader(flags, from_env.split(','), embedding_model, flags.device)
    obs, action, reward, done = data['obs'], data['action'], data['reward'], data['done']
    assert len(obs) == len(action) == len(reward) == len(done), 'data length does not match'
    n_samples = len(reward)
    assert n_samples > 0, 'no data found'
    print('  ', 'total number of samples', n_samples)

    flags.env = to_env
    if make_environment is not None:
        env = make_environment(flags, embedding_model) if mode != 'finetune' else \
            make_environment(flags, embedding_model=None)
    else:
        env = _DatasetEnv(tuple(obs.shape[1:]))
    obs_shape = env.gym_env.observation_space.shape
    policy_cls = PolicyNetWithConv if mode == 'finetune' else PolicyNet
    actor_model = policy_cls(obs_shape, env.gym_env.action_space.n, flags.batch_norm).to(device=flags.device)

    kw = dict(learning_rate=flags.learning_rate, alpha=flags.alpha, epsilon=flags.epsilon, momentum=flags.momentum,
              max_grad_norm=flags.max_grad_norm)
    kw.update(trainer_kwargs or {})
    trainer = BCTrainer(actor_model, obs, action, done, flags.batch_size, flags.unroll_length, flags.max_frames, **kw)
    optimizer, scheduler = trainer.optimizer, trainer.scheduler
    max_epochs = trainer.max_epochs

    if resume:  # main_bc_2.py:93-98
        checkpoint = torch.load(save_path + '.tar', map_location='cpu', weights_only=False)
        if embedding_model is not None:
            embedding_model.load_state_dict(checkpoint["embedding_model_state_dict"])
        actor_model.load_state_dict(checkpoint["actor_model_state_dict"])
        optimizer.load_state_dict(checkpoint["actor_model_optimizer_state_dict"])
        scheduler.load_state_dict(checkpoint["scheduler_state_dict"])

    test_model = None
    if test is not None:
        test_model = policy_cls(obs_shape, env.gym_env.action_space.n, flags.batch_norm).to(device=flags.device)
        test_model.load_state_dict(actor_model.state_dict())
        test_model.eval()

    def evaluate(essential=True):
        if test is None or not essential:
            return {k: np.nan for k in STAT_KEYS}
        test_model.load_state_dict(actor_model.state_dict())
        stats_ep = test(test_model, env, STAT_KEYS, flags.n_episodes_test)
        return {k: np.mean(stats_ep[k]) for k in STAT_KEYS}

    print('=== BC run ===')
    print('  ', 'embedding:', tag)
    print('  ', 'training environment(s):', from_env)
    print('  ', 'testing environment(s):', to_env)
    if resume:
        print('=== Resuming previous run ===')
        for k in ('frames', 'training_loss', 'gradient_norm'):
            print('  ', k.replace('_', ' '), stats[to_env][k][-1])
        init_frames = stats[to_env]['frames'][-1]
    else:
        print('=== Initial evaluation ===')
        stats = {to_env: {**{k: [] for k in STAT_KEYS}, 'frames': [], 'training_loss': [], 'gradient_norm': []}}
        for k, mu in evaluate().items():
            print('  ', k, mu)
            stats[to_env][k].append(mu)
        stats[to_env]['frames'].append(0)
        stats[to_env]['training_loss'].append(np.nan)
        stats[to_env]['gradient_norm'].append(np.nan)
        init_frames = 0
    trainer.frames = init_frames

    print('=== Training policy ===')
    step_frames = flags.batch_size * flags.unroll_length
    for frames in range(init_frames, flags.max_frames, step_frames):
        epoch = frames // step_frames
        loss = trainer.step()
        if (epoch + 1) % flags.eval_frequency == 0:  # main_bc_2.py:230-260
            essential = (not flags.essential_save_only) or is_essential_save(epoch, max_epochs, flags.eval_frequency)
            for k, mu in evaluate(essential).items():
                if essential and test is not None:
                    print('  ', k, mu)
                stats[to_env][k].append(mu)
            loss_v, norm_v = float(loss.item()), float(trainer.gradient_norm().item())
            stats[to_env]['frames'].append(frames)
            stats[to_env]['training_loss'].append(loss_v)
            stats[to_env]['gradient_norm'].append(norm_v)
            print('  ', 'frames', frames)
            print('  ', 'training loss', loss_v)
            print('  ', 'gradient norm', norm_v)
            if not flags.disable_save and _is_writer():
                save_checkpoint(save_path, embedding_model, actor_model, optimizer, scheduler, flags, stats)
    env.close()
    return stats
