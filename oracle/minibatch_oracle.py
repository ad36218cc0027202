"""Test infrastructure (CPU restatement, never on the product path): what make_dataset_naive
(omni_drones/learning/mappo.py:493-513, seq_len == 1) yields for given permutation indices - the rows
``x.reshape(E * T, ...)[indices]`` of every [E, T, ...] tensor of the rollout batch.  Pinned against the reference's own
function by tests/golden/minibatch.npz (oracle/gen_minibatch_golden.py)."""
import numpy as np


def gather(x: np.ndarray, indices: np.ndarray) -> np.ndarray:
    E, T = x.shape[:2]
    return np.ascontiguousarray(x).reshape((E * T,) + x.shape[2:])[indices]


def minibatches(batch: dict, perm: np.ndarray, num_minibatches: int):
    """perm: the reference's randperm over the first (E*T // M) * M samples."""
    for indices in perm.reshape(num_minibatches, -1):
        yield {k: gather(v, indices) for k, v in batch.items()}


def minibatches_seq(batch: dict, perm: np.ndarray, num_minibatches: int, seq_len: int):
    """seq_len > 1 branch (mappo.py:496-505): samples = (env, chunk of seq_len consecutive steps)."""
    out = []
    for indices in perm.reshape(num_minibatches, -1):
        mb = {}
        for k, v in batch.items():
            E, T = v.shape[:2]
            Tc = (T // seq_len) * seq_len
            flat = np.ascontiguousarray(v[:, :Tc]).reshape((E * (Tc // seq_len), seq_len) + v.shape[2:])
            mb[k] = flat[indices]
        out.append(mb)
    return out
