"""TEST INFRASTRUCTURE ONLY - generates tests/golden/policy.npz by running the reference's OWN
modules (loaded by path from /root/reference, build container only, with `tensordict` /
`torchrl.data` stubbed by plain dict classes - SplitEmbedding only iterates the spec and indexes
the tensordict): PartialAttentionEncoder (modules/networks.py:249-314) + DiagGaussian
(modules/distributions.py:66-82) as make_ppo_actor / make_critic assemble them
(mappo.py:575-603, 605-635), on observations shaped like HideAndSeek's
(hideandseek.py:337-349: state_self [1, 20+3F], state_others [A-1, 3], cylinders [K, 5]).
The reference's Actor samples with torch's generator; the golden file stores the noise it drew
(eps = (action - mean) / std) so the restatement and the kernel can be fed the same noise.

Run:  python -m oracle.gen_policy_golden
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/omni_drones/learning/modules/"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "policy.npz")


class Spec:
    def __init__(self, shape):
        self.shape = torch.Size(shape)


class CompositeSpec(dict):
    pass


def load_ref():
    td = types.ModuleType("tensordict"); td.TensorDict = dict
    tr = types.ModuleType("torchrl"); trd = types.ModuleType("torchrl.data")
    trd.CompositeSpec, trd.TensorSpec = CompositeSpec, Spec
    tr.data = trd
    # distributions.py:137 imports TanhNormal for a class this path never builds (cfg actor.tanh: false)
    trm = types.ModuleType("torchrl.modules"); trmd = types.ModuleType("torchrl.modules.distributions")
    trmd.TanhNormal = object
    trm.distributions = trmd
    tr.modules = trm
    stubs = {"tensordict": td, "torchrl": tr, "torchrl.data": trd, "torchrl.modules": trm,
             "torchrl.modules.distributions": trmd}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mods = []
        for name in ("networks", "distributions"):
            spec = importlib.util.spec_from_file_location("_ref_" + name, REF + name + ".py")
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods.append(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods


def case(nets, dists, name, R, D, n_others, K, head_dim, seed, masked_frac=0.3):
    torch.manual_seed(seed)
    spec = CompositeSpec({"state_self": Spec((1, D)), "state_others": Spec((n_others, 3)), "cylinders": Spec((K, 5))})
    enc = nets.PartialAttentionEncoder(spec)
    # default init leaves biases/LayerNorm at 0/1: perturb so that every term is exercised
    with torch.no_grad():
        for n_, q in enc.named_parameters():
            if q.dim() == 1:
                q.add_(0.1 * torch.randn_like(q))
    obs = {"state_self": torch.randn(R, 1, D), "state_others": torch.randn(R, n_others, 3) * 0.5,
           "cylinders": torch.randn(R, K, 5) * 0.5}
    m = torch.rand(R, K) < masked_frac                         # inactive cylinders are rows of -5 (hideandseek.py:769-778)
    obs["cylinders"][m] = -5.0
    obs["state_self"][torch.rand(R) < 0.2, :, :3] = -5.0       # undetected target
    out = {f"{name}/obs/{k}": v.numpy() for k, v in obs.items()}
    with torch.no_grad():
        feat = enc(obs)
        out[f"{name}/ref_feat"] = feat.numpy()
        if head_dim > 1:                                       # actor: DiagGaussian, gain 0.01 like make_ppo_actor
            dist_mod = dists.DiagGaussian(128, head_dim, False, 0.01)
            dist_mod.log_std.add_(0.3 * torch.randn(head_dim))
            dist_mod.fc_mean.bias.add_(0.1 * torch.randn(head_dim))
            dist = dist_mod(feat)
            action = dist.sample()
            out[f"{name}/ref_action"] = action.numpy()
            out[f"{name}/ref_logp"] = dist.log_prob(action).unsqueeze(-1).numpy()
            out[f"{name}/ref_mean"] = dist.mean.numpy()
            out[f"{name}/eps"] = ((action - dist.mean) / dist.stddev).numpy()
            out[f"{name}/ref_logp_mode"] = dist.log_prob(dist.mode).unsqueeze(-1).numpy()
            hw, hb = dist_mod.fc_mean.weight, dist_mod.fc_mean.bias
            out[f"{name}/p/log_std"] = dist_mod.log_std.numpy().copy()
        else:                                                  # critic: v_out (mappo.py:598-603)
            v_out = torch.nn.Linear(128, 1)
            torch.nn.init.orthogonal_(v_out.weight, 0.01)
            out[f"{name}/ref_value"] = v_out(feat).numpy()
            hw, hb = v_out.weight, v_out.bias
        out[f"{name}/p/head.weight"], out[f"{name}/p/head.bias"] = hw.numpy().copy(), hb.numpy().copy()
        for k, v in enc.state_dict().items():
            out[f"{name}/p/{k}"] = v.numpy().copy()
    return out


def main():
    nets, dists = load_ref()
    data = {}
    data.update(case(nets, dists, "actor_tp", 200, 35, 2, 3, 4, 1))        # HideAndSeek defaults, use_TP_net=1
    data.update(case(nets, dists, "critic_tp", 200, 35, 2, 3, 1, 2))
    data.update(case(nets, dists, "actor_notp", 77, 20, 2, 3, 4, 3))       # use_TP_net=0: state_self width 20
    data.update(case(nets, dists, "actor_a2_k2", 33, 35, 1, 2, 4, 4))      # 2 pursuers, 2 observed cylinders
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
