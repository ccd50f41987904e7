"""BC sampling helpers with the reference's semantics (src/utils_bc.py:5-29): same draws from Python's `random`."""
import random

import numpy as np


def is_essential_save(epoch, max_epochs, eval_frequency):
    """True inside +-5 evaluation periods around 1 %, 10 %, 50 % and 97 % of training (src/utils_bc.py:5-12)."""
    window = 5 * eval_frequency
    return any(int(frac * max_epochs) - window <= epoch < int(frac * max_epochs) + window
               for frac in (0.01, 0.1, 0.5, 0.97))


def sample_with_minimum_distance(n=40, k=4, d=10):
    """k start indices in range(n), pairwise at least d apart (src/utils_bc.py:24-29): draw k distinct values from
    the range shrunk by (k-1)(d-1) and push the r-th smallest up by r(d-1). Order of the draw is preserved."""
    draw = random.sample(range(n - (k - 1) * (d - 1)), k)
    rank = np.empty(k, dtype=np.int64)
    rank[np.argsort(np.asarray(draw), kind="stable")] = np.arange(k)
    return [int(s + (d - 1) * r) for s, r in zip(draw, rank)]


def window_indices(starting_i, unroll_length, n_samples):
    """(T, B) sample indices of the B windows, wrapping modulo n_samples (main_bc_2.py:194-201)."""
    start = np.asarray(starting_i, dtype=np.int64)[None, :]
    return np.mod(start + np.arange(unroll_length, dtype=np.int64)[:, None], n_samples)
