"""Policy inference next to the tick: the torch-CPU restatement
oracle/policy_oracle.py against tests/golden/policy.npz - produced by the reference's own
PartialAttentionEncoder / DiagGaussian modules (oracle/gen_policy_golden.py) - and, on the GPU, the
fused kernel hs_policy_forward (through mupe_b200.FusedPolicy) against both.  Tolerance: 1e-4
relative fp32 (north_star), absolute floor 2e-5 on O(1) LayerNorm outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import policy_oracle as PO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "policy.npz")
CASES = ["actor_tp", "critic_tp", "actor_notp", "actor_a2_k2"]


def _load(name):
    z = np.load(GOLD)
    p = {k.split("/p/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/p/")}
    obs = {k.split("/obs/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/obs/")}
    ref = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files
           if k.startswith(name + "/ref_") or k == name + "/eps"}
    return p, obs, ref


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_modules(name):
    p, obs, ref = _load(name)
    torch.testing.assert_close(PO.encoder(p, obs), ref["ref_feat"], rtol=1e-5, atol=5e-6)
    if "log_std" in p:
        a, lp, mean = PO.actor(p, obs, ref["eps"])
        torch.testing.assert_close(mean, ref["ref_mean"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(a, ref["ref_action"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lp, ref["ref_logp"], rtol=1e-5, atol=1e-5)
        _, lp0, _ = PO.actor(p, obs, None)
        torch.testing.assert_close(lp0, ref["ref_logp_mode"], rtol=1e-5, atol=1e-5)
    else:
        torch.testing.assert_close(PO.critic(p, obs), ref["ref_value"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_policy_matches_reference_modules(built_lib, name):
    import mupe_b200
    dev = torch.device("cuda:0")
    p, obs, ref = _load(name)
    pg = {k: v.to(dev).contiguous() for k, v in p.items()}
    og = {k: v.to(dev).contiguous() for k, v in obs.items()}
    net = mupe_b200.FusedPolicy(pg, n_others=obs["state_others"].shape[1], n_cyl=obs["cylinders"].shape[1], device=dev)
    is_actor = "log_std" in p
    eps = ref["eps"].to(dev).contiguous() if is_actor else None
    out = net(og["state_self"], og["state_others"], og["cylinders"], eps=eps, want_features=True)
    torch.testing.assert_close(out["features"].cpu(), ref["ref_feat"], rtol=1e-4, atol=2e-5)
    if is_actor:
        torch.testing.assert_close(out["head"].cpu(), ref["ref_mean"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["action"].cpu(), ref["ref_action"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["logp"].cpu(), ref["ref_logp"], rtol=1e-4, atol=1e-5)
        mode = net(og["state_self"], og["state_others"], og["cylinders"], eps=None)
        torch.testing.assert_close(mode["action"].cpu(), ref["ref_mean"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(mode["logp"].cpu(), ref["ref_logp_mode"], rtol=1e-4, atol=1e-5)
    else:
        torch.testing.assert_close(out["head"].cpu(), ref["ref_value"], rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 31, 33, 4096 * 3, 148 * 128 + 5])
def test_gpu_policy_ragged_sizes_against_oracle(built_lib, R):
    """Row counts around the 32- and 64-row tiles (both kernels), against the restatement on seeded inputs; refresh()
    picks up in-place parameter updates."""
    import mupe_b200
    dev = torch.device("cuda:0")
    p, _, _ = _load("actor_tp")
    g = torch.Generator().manual_seed(R)
    obs = {"state_self": torch.randn(R, 1, 35, generator=g), "state_others": torch.randn(R, 2, 3, generator=g),
           "cylinders": torch.randn(R, 3, 5, generator=g)}
    eps = torch.randn(R, 4, generator=g)
    pg = {k: v.to(dev).contiguous() for k, v in p.items()}
    net = mupe_b200.FusedPolicy(pg, 2, 3, dev)
    og = {k: v.to(dev) for k, v in obs.items()}
    out = net(og["state_self"], og["state_others"], og["cylinders"], eps=eps.to(dev))
    n = min(R, 3000)                      # the CPU restatement on a prefix and a suffix is enough at the large sizes
    for sl in (slice(0, n), slice(R - n, R)):
        a, lp, mean = PO.actor(p, {k: v[sl] for k, v in obs.items()}, eps[sl])
        torch.testing.assert_close(out["action"][sl].cpu(), a, rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["logp"][sl].cpu(), lp, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        pg["head.bias"].add_(1.0)
    out2 = net.refresh()(og["state_self"], og["state_others"], og["cylinders"], eps=eps.to(dev))
    torch.testing.assert_close(out2["head"], out["head"] + 1.0, rtol=1e-5, atol=1e-5)
