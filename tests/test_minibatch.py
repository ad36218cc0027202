"""PPO minibatches straight from the rollout: hs_gather_rows / make_dataset_naive against
tests/golden/minibatch.npz, which oracle/gen_minibatch_golden.py produces by running the reference's OWN
make_dataset_naive (omni_drones/learning/mappo.py:493-513)."""
import os

import numpy as np
import pytest
import torch

from oracle import minibatch_oracle as MO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "minibatch.npz")


def _load():
    z = np.load(GOLD)
    batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    return z, batch, int(z["E"]), int(z["T"]), int(z["M"])


def test_oracle_matches_the_reference_minibatches():
    z, batch, E, T, M = _load()
    for m, mb in enumerate(MO.minibatches(batch, z["perm"], M)):
        for k, v in mb.items():
            assert np.array_equal(v, z[f"mb{m}_{k}"]), (m, k)


def test_oracle_matches_the_reference_minibatches_of_step_chunks():
    z, batch, E, T, M = _load()
    for m, mb in enumerate(MO.minibatches_seq(batch, z["perm_seq2"], M, 2)):
        for k, v in mb.items():
            assert np.array_equal(v, z[f"seq2_mb{m}_{k}"]), (m, k)


@pytest.mark.gpu
def test_gather_matches_the_reference_minibatches_of_step_chunks():
    """seq_len = 2 (mappo.py:496-505): samples of two consecutive steps, from time-major storage views."""
    import mupe_b200
    z, batch, E, T, M = _load()
    dev = torch.device("cuda:0")
    tb = {k: torch.from_numpy(v).to(dev).transpose(0, 1).contiguous().transpose(0, 1) for k, v in batch.items()}
    got = list(mupe_b200.make_dataset_naive(tb, M, seq_len=2, perm=torch.from_numpy(z["perm_seq2"])))
    assert len(got) == M
    for m, mb in enumerate(got):
        for k in batch:
            assert np.array_equal(mb[(k,)].cpu().numpy(), z[f"seq2_mb{m}_{k}"]), (m, k)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["env_major", "time_major"])
def test_gather_matches_the_reference_minibatches(layout):
    """Bit-exact, from the reference's [E, T, ...] layout and from [E, T] views of time-major [T, E, ...] storage."""
    import mupe_b200
    z, batch, E, T, M = _load()
    dev = torch.device("cuda:0")
    tb = {}
    for k, v in batch.items():
        t = torch.from_numpy(v).to(dev)
        tb[k] = t if layout == "env_major" else t.transpose(0, 1).contiguous().transpose(0, 1)
        assert tb[k].is_contiguous() == (layout == "env_major")
    got = list(mupe_b200.make_dataset_naive({"agents": {"obs": tb["state_self"]}, **{k: v for k, v in tb.items() if k != "state_self"}},
                                            M, perm=torch.from_numpy(z["perm"])))
    assert len(got) == M
    for m, mb in enumerate(got):
        assert np.array_equal(mb[("agents", "obs")].cpu().numpy(), z[f"mb{m}_state_self"])
        for k in batch:
            if k != "state_self":
                assert np.array_equal(mb[(k,)].cpu().numpy(), z[f"mb{m}_{k}"]), (m, k)
                assert mb[(k,)].dtype == tb[k].dtype and mb[(k,)].is_contiguous()


@pytest.mark.gpu
def test_gather_full_size_from_the_rollout_storage():
    """4096 x 64 rollout rows of an engine's time-major storage: every minibatch equals torch's own indexing of the
    flattened copy, the minibatches partition the first (E*T // M) * M samples, and a second draw differs."""
    import mupe_b200
    dev = torch.device("cuda:0")
    E, T, M = 4096, 64, 16
    eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev, rollout_steps=T)
    b = eng.storage.batch()
    for k in ("state_self", "reward", "tp_input"):
        b[k].copy_(torch.randn(b[k].shape, device=dev))
    sel = {k: b[k] for k in ("state_self", "state_others", "obs_cylinders", "reward", "done", "tp_input")}
    torch.manual_seed(0)
    perm = torch.randperm(E * T, device=dev)
    flat = {k: v.reshape(E * T, *v.shape[2:]) for k, v in sel.items()}          # the copy the reference makes
    seen = []
    for m, mb in enumerate(mupe_b200.make_dataset_naive(sel, M, perm=perm)):
        idx = perm.reshape(M, -1)[m]
        seen.append(idx)
        for k in sel:
            assert torch.equal(mb[(k,)], flat[k][idx]), (m, k)
    assert torch.equal(torch.cat(seen).sort().values, torch.arange(E * T, device=dev))
    a = next(iter(mupe_b200.make_dataset_naive(sel, M)))[("reward",)]
    c = next(iter(mupe_b200.make_dataset_naive(sel, M)))[("reward",)]
    assert not torch.equal(a, c)
    mb8 = next(iter(mupe_b200.make_dataset_naive(sel, M, seq_len=8)))
    assert tuple(mb8[("reward",)].shape) == (E * (T // 8) // M, 8, 3, 1)
    eng.close()
