"""TEST INFRASTRUCTURE ONLY (never imported by the product path) - CPU restatement of the
advantage computation the reference runs on a finished rollout:

* compute_gae            omni_drones/learning/utils/gae.py:27-51
* batch normalisation    omni_drones/learning/mappo.py:391-396  ((adv - mean) / (std + 1e-8),
                         torch.Tensor.std() = Bessel-corrected)

numpy fp32 with one rounding per operation, in the reference's operation order.  Pinned against
the reference's own gae.py (imported from /root/reference by oracle/gen_gae_golden.py) through
tests/golden/gae.npz: bit-exact advantages and returns.
"""
import numpy as np


def compute_gae(reward, done, value, next_value, gamma=0.99, lmbda=0.95):
    """reward, value [N,T,k] fp32; done [N,T,1] bool/uint8; next_value [N,k] -> advantages, returns [N,T,k]."""
    reward = np.asarray(reward, np.float32)
    value = np.asarray(value, np.float32)
    assert reward.shape == value.shape
    f = np.float32
    not_done = (f(1.0) - np.asarray(done).astype(np.float32)).astype(np.float32)        # gae.py:37
    T = reward.shape[1]
    nv = np.asarray(next_value, np.float32)
    g, gl = f(gamma), f(gamma * lmbda)            # python-double product, rounded once (scalar * tensor)
    gae = np.zeros_like(nv)
    adv = np.zeros_like(reward)
    for t in reversed(range(T)):
        nd = not_done[:, t]
        delta = ((reward[:, t] + ((g * nv).astype(f) * nd).astype(f)).astype(f) - value[:, t]).astype(f)   # :41-45
        gae = (delta + ((gl * nd).astype(f) * gae).astype(f)).astype(f)                                     # :46
        adv[:, t] = gae
        nv = value[:, t]
    return adv, (adv + value).astype(f)           # :49


def normalize_advantages(adv):
    """mappo.py:391-396; float64 moments (torch reduces in fp32 with a cascade, agreement ~1e-6)."""
    a = np.asarray(adv, np.float64)
    mean, std = a.mean(), a.std(ddof=1)
    out = ((np.asarray(adv, np.float32) - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))).astype(np.float32)
    return out, np.float32(mean), np.float32(std)
