"""Offline embedding of a trajectory dataset — drop-in for behavioral_cloning/save_embedded_obs.py (`run(flags)` with the
same flags, file names and pickle layouts), on top of the fused CUDA path.

Artefacts (same as the reference, save_embedded_obs.py:95-172):
  <data_path>/<embedding>[_<run_id>].tar      {'embedding_model_state_dict': EmbeddingNet.state_dict()}   (:125-131)
  <data_path>/<env>_<embedding>.pickle         dict(obs (N, O*n) float32, action, reward, done, true_state) flat over
                                               steps (+ 'png': list of frame paths for --source png)      (:159-172)
and nothing is recomputed if the pickle already exists (:99-100).

What changes is how the observations get through the encoder. The reference loops over mini-batches of `batch_size`
observations, splits frames on the host, calls the encoder, reads the result back and regroups it
(save_embedded_obs.py:146-156); here whole passes of observations go through `EmbeddingNet.embed(obs, n_frames)` — the
frame split, preprocessing, encoder and regrouping run on the GPU and frame f of sample i lands in
out[i, f*O:(f+1)*O] — and the rows are copied into one preallocated host array. With --source png the reference embeds
one frame per call (:50-93); here the goal image and all frames of a trajectory go through the encoder as one batch
(the kernels' results do not depend on the batch composition). Under torchrun (`torch.distributed` initialised) the
samples are cut into contiguous blocks per rank (no data-path collective), the blocks are gathered on rank 0, and rank 0
alone writes the files.
"""
import os
import pickle
import random

import numpy as np
import torch

from . import parallel
from .embeddings import EmbeddingNet


def read_habitat_data_from_pickle(data_path, n_trajectories=-1):
    """save_embedded_obs.py:29-48: lists over trajectories -> arrays over steps."""
    print('loading %s ...' % data_path)
    with open(data_path + '.pickle', 'rb') as fh:
        data = pickle.load(fh)
    if n_trajectories == -1:
        n_trajectories = len(data['reward'])
    for k in ('obs', 'action', 'reward', 'done', 'true_state'):
        data[k] = np.concatenate(data[k][:n_trajectories])
    n_samples = len(data['reward'])
    print('  ', '%d trajectories for a total of %d samples' % (n_trajectories, n_samples))
    print('  ', 'avg. return is', data['reward'].sum() / n_trajectories)
    return data


def _rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def embed_observations(embedding_model, obs, pass_size=4096):
    """(N, H, W, 3n) uint8 array -> (N, O*n) float32 host array; this rank's block only under torch.distributed
    (see `gather_rows`). Passes of `pass_size` observations bound the device memory taken by the input frames."""
    n_samples = obs.shape[0]
    n_frames = max(obs.shape[3] // 3, 1)
    rank, world = _rank_world()
    lo, hi = parallel.shard_range(n_samples, rank, world)
    out = np.empty((hi - lo, embedding_model.out_size * n_frames), dtype=np.float32)
    for i in range(lo, hi, pass_size):
        j = min(i + pass_size, hi)
        o = embedding_model.embed(torch.from_numpy(np.ascontiguousarray(obs[i:j])), n_frames)
        out[i - lo:j - lo] = o.detach().cpu().numpy() if isinstance(o, torch.Tensor) else np.asarray(o)
    return out


def gather_rows(block, n_samples):
    """Concatenate the per-rank blocks of `embed_observations` on rank 0 (None elsewhere); identity for one process."""
    rank, world = _rank_world()
    if world == 1:
        return block
    width = block.shape[1]
    rows = -(-n_samples // world)  # blocks differ by at most one row: pad to the largest
    backend = torch.distributed.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.zeros(rows, width, dtype=torch.float32, device=dev)
    mine[:block.shape[0]] = torch.from_numpy(block).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    torch.distributed.gather(mine, parts, dst=0)
    if rank != 0:
        return None
    out = np.empty((n_samples, width), dtype=np.float32)
    for r in range(world):
        lo, hi = parallel.shard_range(n_samples, r, world)
        out[lo:hi] = parts[r][:hi - lo].cpu().numpy()
    return out


def read_habitat_data_from_png(data_path, model=None, n_trajectories=-1):
    """save_embedded_obs.py:50-93. `<t>.pickle` holds the trajectory's action / reward / done / true_state,
    `<t>_goal.png` the goal image and `<t>_<s>.png` the frames, read with cv2 in file channel order like the
    reference. Each sample is [embedding(frame) | embedding(goal)] (raw pixels when `model` is None)."""
    import cv2
    print('loading %s ...' % data_path)
    data = dict(obs=[], action=[], reward=[], done=[], true_state=[], png=[])
    if n_trajectories == -1:
        n_trajectories = 100000
    t = 0
    for t in range(n_trajectories):
        meta = os.path.join(data_path, str(t) + '.pickle')
        goal = cv2.imread(os.path.join(data_path, str(t) + '_goal.png')) if os.path.isfile(meta) else None
        if goal is None:
            break
        with open(meta, 'rb') as fh:
            tmp = pickle.load(fh)
        for k in data.keys():
            if k in tmp:
                data[k].append(tmp[k])
        frames, names = [], []
        for s in range(500):  # 500 is the max step per trajectory according to Habitat's YAML config
            name = os.path.join(data_path, str(t) + '_' + str(s)) + '.png'
            img = cv2.imread(name)
            if img is None:
                break
            frames.append(img)
            names.append(name)
        if model is not None and frames:
            emb = model.embed(torch.from_numpy(np.stack([goal] + frames)), 1)
            emb = emb.detach().cpu().numpy() if isinstance(emb, torch.Tensor) else np.asarray(emb)
            goal_e, frames_e = emb[0], emb[1:]
        else:
            goal_e, frames_e = goal, frames
        for f in frames_e:
            data['obs'].append(np.concatenate((f, goal_e), -1))
        data['png'] += names
    n_trajectories = t
    data['obs'] = np.stack(data['obs'])
    for k in ('action', 'reward', 'done', 'true_state'):
        data[k] = np.concatenate(data[k])
    n_samples = len(data['reward'])
    print('  ', '%d trajectories for a total of %d samples' % (n_trajectories, n_samples))
    print('  ', 'avg. return is', data['reward'].sum() / max(n_trajectories, 1))
    return data


def run(flags, embedding_factory=EmbeddingNet):
    """`flags`: the reference's argparse namespace (src/arguments.py + --n_trajectories, --source). `embedding_factory`
    exists for the host-logic tests (an object with `out_size`, `state_dict()` and `embed(obs, n_frames)`)."""
    save_name = os.path.join(flags.data_path, flags.env + '_' + flags.embedding_name + '.pickle')
    if os.path.isfile(save_name):
        return
    rank, world = _rank_world()

    # Fix seeds (save_embedded_obs.py:102-106)
    torch.manual_seed(flags.run_id)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(flags.run_id)
    np.random.seed(flags.run_id)
    random.seed(flags.run_id)

    embedding_model = embedding_factory(flags.embedding_name, in_channels=3, pretrained=flags.pretrained_embedding,
                                        train=flags.train_embedding, disable_cuda=flags.disable_cuda)

    # Save model that will be used in main_bc
    emb_path = os.path.join(flags.data_path, flags.embedding_name)
    if flags.embedding_name == 'random':
        emb_path += '_' + str(flags.run_id)
    if rank == 0:
        torch.save({'embedding_model_state_dict': embedding_model.state_dict()}, emb_path + '.tar')

    print('=== Loading trajectories ===')
    source = getattr(flags, 'source', 'png')
    if source == 'png':
        if world > 1:
            raise NotImplementedError("--source png is read trajectory by trajectory on one process")
        data = read_habitat_data_from_png(os.path.join(flags.data_path, flags.env), embedding_model,
                                          getattr(flags, 'n_trajectories', -1))
    else:
        data = read_habitat_data_from_pickle(os.path.join(flags.data_path, flags.env))
        print('  ', 'passing observations through embedding model')
        n_samples = data['obs'].shape[0]
        obs = gather_rows(embed_observations(embedding_model, data['obs']), n_samples)
        data = dict(obs=obs, action=data['action'][:n_samples], reward=data['reward'][:n_samples],
                    done=data['done'][:n_samples], true_state=data['true_state'][:n_samples])
    if rank != 0:
        return
    n_samples = len(data['reward'])
    assert n_samples > 0, 'no data found'
    print('  ', 'total number of samples', n_samples)
    with open(save_name, 'wb') as handle:
        pickle.dump(data, handle, protocol=pickle.HIGHEST_PROTOCOL)
