"""Command-line flags of the BC / embedding entry scripts — the names, types and defaults of the reference's shared
parser (src/arguments.py:3-68), which its scripts import and extend (behavioral_cloning/save_embedded_obs.py:25-26).
`make_parser()` returns a fresh parser (the reference keeps one module-level instance: `parser` below mirrors that)."""
import argparse


def make_parser():
    p = argparse.ArgumentParser(description='pvr_habitat_b200 behavioural cloning')
    add = p.add_argument
    # behavioural cloning (src/arguments.py:5-14)
    add('--max_frames', type=int, default=200000000)
    add('--n_episodes_test', type=int, default=50)
    add('--eval_frequency', type=int, default=200)
    add('--to_env', type=str, default='HabitatImageNav-apartment_0')
    add('--debug', action='store_true')
    add('--disable_save', action='store_true')
    add('--essential_save_only', action='store_true')
    add('--save_path', type=str, default='bc')
    add('--data_path', type=str, default='behavioral_cloning')
    # embedding (:17-24)
    add('--embedding_name', type=str, default='resnet50', help='Name of the embedding model.')
    add('--train_embedding', action='store_true', help='Train observation embedding or keep it fixed.')
    add('--disable_pretrained_embedding', action='store_false', dest='pretrained_embedding',
        help='Use it to prevent loading pretrained weights.')
    add('--batch_norm', action='store_true', help='Place a BatchNorm1d layer at the beginning of the policy.')
    # environment (:27-34)
    add('--env', type=str, default='HabitatImageNav-apartment_0',
        help='Training environments; several, trained together, as a comma-separated list.')
    add('--num_input_frames', type=int, default=1, help='Number of input frames per observation.')
    # general (:37-44)
    add('--xpid', default=None, help='Experiment ID.')
    add('--run_id', default=1, type=int, help='Run ID (doubles as the random seed of the BC scripts).')
    add('--seed', default=1, type=int, help='Random seed.')
    # training (:47-57)
    add('--total_frames', default=50000000, type=int, help='Total environment frames to train for.')
    add('--batch_size', default=32, type=int, help='Learner batch size.')
    add('--unroll_length', default=100, type=int, help='The unroll length (time dimension).')
    add('--mp_start', default='spawn', type=str, help='Start method of multiprocesses (unused, as in the reference).')
    add('--disable_cuda', action='store_true', help='Disable CUDA (rejected: there is no CPU path).')
    # optimizer (:60-68)
    add('--learning_rate', default=0.0001, type=float, help='Learning rate.')
    add('--alpha', default=0.99, type=float, help='RMSProp smoothing constant.')
    add('--momentum', default=0, type=float, help='RMSProp momentum.')
    add('--epsilon', default=1e-5, type=float, help='RMSProp epsilon.')
    add('--max_grad_norm', default=40., type=float, help='Max norm of gradients.')
    return p


parser = make_parser()
