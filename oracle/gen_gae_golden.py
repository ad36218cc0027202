"""TEST INFRASTRUCTURE ONLY - generates tests/golden/gae.npz by running the reference's OWN
omni_drones/learning/utils/gae.py (pure torch; loaded by path from /root/reference, build
container only) on seeded rollouts shaped like MAPPOPolicy.train_op's call
(omni_drones/learning/mappo.py:381-397: reward/value [E,T,A,1], done [E,T,A,1] = env done
broadcast over the agents, next_value [E,A,1], gamma/lambda of cfg/algo/mappo.yaml), followed by
the batch normalisation of mappo.py:391-396.

Run:  python -m oracle.gen_gae_golden
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/omni_drones/learning/utils/gae.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gae.npz")


def load_ref():
    spec = importlib.util.spec_from_file_location("_ref_gae", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def case(ref, name, E, T, A, p_done, gamma, lmbda, seed):
    g = torch.Generator().manual_seed(seed)
    reward = torch.randn(E, T, A, 1, generator=g) * 3.0
    # catch reward spikes and collision penalties give the rewards a long tail
    reward += (torch.rand(E, T, A, 1, generator=g) < 0.02).float() * 20.0 - (torch.rand(E, T, A, 1, generator=g) < 0.01).float() * 100.0
    value = torch.randn(E, T, A, 1, generator=g) * 5.0
    next_value = torch.randn(E, A, 1, generator=g) * 5.0
    env_done = torch.rand(E, T, 1, generator=g) < p_done
    done = env_done.unsqueeze(-1).expand(E, T, A, 1)                 # MAPPOPolicy._get_dones, mappo.py:358-365
    adv, ret = ref.compute_gae(reward, done, value, next_value, gamma=gamma, lmbda=lmbda)
    mean, std = adv.mean(), adv.std()
    adv_n = (adv - mean) / (std + 1e-8)
    return {f"{name}/reward": reward.numpy(), f"{name}/value": value.numpy(), f"{name}/next_value": next_value.numpy(),
            f"{name}/done": env_done.numpy(), f"{name}/gamma": np.float64(gamma), f"{name}/lmbda": np.float64(lmbda),
            f"{name}/ref_adv": adv.numpy(), f"{name}/ref_ret": ret.numpy(), f"{name}/ref_adv_norm": adv_n.numpy(),
            f"{name}/ref_mean": mean.numpy(), f"{name}/ref_std": std.numpy()}


def main():
    ref = load_ref()
    data = {}
    data.update(case(ref, "mappo_default", 96, 64, 3, 1.0 / 800, 0.995, 0.95, 1))      # cfg/algo/mappo.yaml
    data.update(case(ref, "many_dones", 40, 33, 3, 0.2, 0.99, 0.95, 2))
    data.update(case(ref, "single_agent_T1", 17, 1, 1, 0.5, 0.9, 0.8, 3))
    data.update(case(ref, "no_dones", 8, 128, 2, 0.0, 0.995, 0.95, 4))
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
