"""Advantage scan: the numpy restatement oracle/gae_oracle.py against
tests/golden/gae.npz - produced by the reference's own omni_drones/learning/utils/gae.py
(oracle/gen_gae_golden.py) - and, on the GPU, hs_gae through the reference-named wrapper
mupe_b200.rollout.compute_gae against both (bit-exact advantages/returns; 1e-5 for the
normalised advantages, whose mean/std are reductions)."""
import os

import numpy as np
import pytest

from oracle import gae_oracle as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gae.npz")
CASES = ["mappo_default", "many_dones", "single_agent_T1", "no_dones"]


def _case(z, name):
    g = lambda k: z[f"{name}/{k}"]
    return dict(reward=g("reward"), value=g("value"), next_value=g("next_value"), done=g("done"),
                gamma=float(g("gamma")), lmbda=float(g("lmbda")), ref_adv=g("ref_adv"), ref_ret=g("ref_ret"),
                ref_adv_norm=g("ref_adv_norm"), ref_mean=float(g("ref_mean")), ref_std=float(g("ref_std")))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_gae(name):
    c = _case(np.load(GOLD), name)
    E, T, A, _ = c["reward"].shape
    done = np.broadcast_to(c["done"][:, :, None, :], (E, T, A, 1))
    adv, ret = G.compute_gae(c["reward"], done, c["value"], c["next_value"], c["gamma"], c["lmbda"])
    assert np.array_equal(adv, c["ref_adv"]), "advantages differ from the reference's compute_gae"
    assert np.array_equal(ret, c["ref_ret"]), "returns differ from the reference's compute_gae"
    if adv.size > 1:
        adv_n, mean, std = G.normalize_advantages(adv)
        assert abs(mean - c["ref_mean"]) <= 1e-5 * max(1.0, abs(c["ref_mean"]))
        assert abs(std - c["ref_std"]) <= 1e-5 * c["ref_std"]
        np.testing.assert_allclose(adv_n, c["ref_adv_norm"], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["env_major", "time_major"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_gae_matches_reference(built_lib, name, layout):
    import torch
    from mupe_b200 import rollout as R
    c = _case(np.load(GOLD), name)
    dev = torch.device("cuda:0")
    E, T, A, _ = c["reward"].shape

    def put(x):          # [E,T,...] tensor whose storage is either [E,T,...] or [T,E,...]
        t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        return t if layout == "env_major" else t.transpose(0, 1).contiguous().transpose(0, 1)
    reward, value, done = put(c["reward"]), put(c["value"]), put(c["done"])
    nv = torch.from_numpy(c["next_value"]).to(dev)
    adv, ret, stats = R.compute_gae(reward, done.unsqueeze(2).expand(E, T, A, 1), value, nv, c["gamma"], c["lmbda"],
                                    return_stats=True)
    assert adv.stride() == reward.stride()
    assert np.array_equal(adv.cpu().numpy(), c["ref_adv"])
    assert np.array_equal(ret.cpu().numpy(), c["ref_ret"])
    if adv.numel() > 1:
        mean, std = stats.cpu().tolist()
        assert abs(mean - c["ref_mean"]) <= 1e-5 * max(1.0, abs(c["ref_mean"]))
        assert abs(std - c["ref_std"]) <= 1e-5 * c["ref_std"]
        adv_n, _ = R.compute_gae(reward, done, value, nv, c["gamma"], c["lmbda"], normalize=True)
        np.testing.assert_allclose(adv_n.cpu().numpy(), c["ref_adv_norm"], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_gpu_gae_full_size_properties(built_lib):
    """65 536 envs x 64 steps x 3 agents: (a) against torch's eager loop on the GPU (the reference's algorithm,
    same device), (b) linearity in the rewards with dones and values fixed, (c) a done step cuts the scan."""
    import torch
    from mupe_b200 import rollout as R
    dev = torch.device("cuda:0")
    E, T, A = 65536, 64, 3
    g = torch.Generator(device=dev).manual_seed(0)
    reward = torch.randn(T, E, A, 1, generator=g, device=dev).transpose(0, 1)
    value = torch.randn(T, E, A, 1, generator=g, device=dev).transpose(0, 1)
    nv = torch.randn(E, A, 1, generator=g, device=dev)
    done = (torch.rand(T, E, 1, generator=g, device=dev) < 0.01).transpose(0, 1)
    gamma, lmbda = 0.995, 0.95
    adv, ret = R.compute_gae(reward, done, value, nv, gamma, lmbda)
    # (a) the reference's loop (gae.py:39-47) in eager torch on the same device
    nd = 1.0 - done.unsqueeze(2).expand(E, T, A, 1).float()
    gae, nxt, want = 0, nv, torch.zeros_like(reward)
    for t in reversed(range(T)):
        delta = reward[:, t] + gamma * nxt * nd[:, t] - value[:, t]
        want[:, t] = gae = delta + (gamma * lmbda * nd[:, t] * gae)
        nxt = value[:, t]
    torch.testing.assert_close(adv, want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ret, want + value, rtol=1e-5, atol=1e-5)
    # (b) with value = 0 and next_value = 0 the map reward -> advantages is linear
    z, nz = torch.zeros_like(value), torch.zeros_like(nv)
    a1, _ = R.compute_gae(reward, done, z, nz, gamma, lmbda)
    a2, _ = R.compute_gae(2.0 * reward, done, z, nz, gamma, lmbda)
    torch.testing.assert_close(a2, 2.0 * a1, rtol=1e-6, atol=1e-6)
    # (c) at a done step the advantage is reward - value (nothing flows across the boundary)
    m = done.unsqueeze(2).expand(E, T, A, 1)
    torch.testing.assert_close(adv[m], (reward - value)[m], rtol=0, atol=0)
