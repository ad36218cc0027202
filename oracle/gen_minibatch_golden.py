"""Generates tests/golden/minibatch.npz by running the REFERENCE'S OWN `make_dataset_naive`
(omni_drones/learning/mappo.py:493-513, AST-extracted; the pinned tensordict fork is not in this image, so the function
is handed a minimal stand-in with the three operations it uses: .shape / .device, .reshape(-1), [indices]).
Test infrastructure; run in the build container:  python -m oracle.gen_minibatch_golden"""
import ast
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ref_harness as RH  # noqa: E402


class MiniTD:
    def __init__(self, d, batch):
        self.d, self.shape = d, torch.Size(batch)

    @property
    def device(self):
        return next(iter(self.d.values())).device

    def reshape(self, *shape):
        n = len(self.shape)
        new = {k: v.reshape(*shape, *v.shape[n:]) for k, v in self.d.items()}
        return MiniTD(new, next(iter(new.values())).shape[:len(shape)])

    def __getitem__(self, idx):
        new = {k: v[idx] for k, v in self.d.items()}
        return MiniTD(new, next(iter(new.values())).shape[:len(self.shape)])


def main():
    src = (RH.REF / "omni_drones/learning/mappo.py").read_text()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "make_dataset_naive")
    ns = dict(torch=torch, TensorDict=MiniTD)
    exec(compile(ast.Module(body=[node], type_ignores=[]), "<ref:make_dataset_naive>", "exec"), ns)
    E, T, A, M = 6, 5, 3, 4
    g = torch.Generator().manual_seed(11)
    batch = {"state_self": torch.randn(E, T, A, 1, 35, generator=g), "reward": torch.randn(E, T, A, 1, generator=g),
             "state_value": torch.randn(E, T, A, 1, generator=g), "done": torch.rand(E, T, 1, generator=g) < 0.3,
             "tp_input": torch.randn(E, T, 10, 16, generator=g), "odd": torch.randn(E, T, 7, generator=g)}
    torch.manual_seed(5)
    state = torch.get_rng_state()
    outs = list(ns["make_dataset_naive"](MiniTD(batch, (E, T)), M, 1))
    torch.set_rng_state(state)
    perm = torch.randperm((E * T // M) * M)          # the draw the function made (same generator state)
    out = {"E": np.asarray(E), "T": np.asarray(T), "M": np.asarray(M), "perm": perm.numpy()}
    for k, v in batch.items():
        out["in_" + k] = v.numpy()
    for m, td in enumerate(outs):
        for k, v in td.d.items():
            out[f"mb{m}_{k}"] = v.numpy()
        # sanity: the recorded permutation reproduces the minibatch
        idx = perm.reshape(M, -1)[m]
        assert torch.equal(batch["reward"].reshape(E * T, A, 1)[idx], td.d["reward"])
    # seq_len = 2: chunks of two consecutive steps (T = 5 -> the last step is dropped), 12 samples, 3 per minibatch
    torch.manual_seed(6)
    state = torch.get_rng_state()
    outs2 = list(ns["make_dataset_naive"](MiniTD(batch, (E, T)), M, 2))
    torch.set_rng_state(state)
    perm2 = torch.randperm((E * (T // 2) // M) * M)
    out["perm_seq2"] = perm2.numpy()
    for m, td in enumerate(outs2):
        for k, v in td.d.items():
            out[f"seq2_mb{m}_{k}"] = v.numpy()
    path = os.path.join(REPO, "tests", "golden", "minibatch.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(outs), "minibatches of", outs[0].shape[0], "samples")


if __name__ == "__main__":
    main()
