"""TEST INFRASTRUCTURE ONLY (never imported by the product path) - CPU restatement of the policy
inference that runs next to the environment tick:

* SplitEmbedding.forward          omni_drones/learning/modules/networks.py:153-161
* PartialAttentionEncoder.forward omni_drones/learning/modules/networks.py:280-314
  (nn.MultiheadAttention, one head, query = token 0, keys/values = all tokens, post-norm)
* DiagGaussian.forward            omni_drones/learning/modules/distributions.py:78-82
* Actor.forward (sample / mode + log_prob), Critic.forward (v_out)
                                  omni_drones/learning/mappo.py:614-635, 652-668

Written with explicit matrix products in torch CPU fp32 (no nn.Module), straight from the
definitions, so that it is an independent statement of the arithmetic.  Pinned against the
reference's own modules through tests/golden/policy.npz (oracle/gen_policy_golden.py).

Parameter dict ``p`` uses the reference's state_dict names relative to the encoder:
  split_embed.embed.{state_self,state_others,cylinders}.{weight,bias}, split_embed.layer_norm.*,
  attn.in_proj_weight/bias, attn.out_proj.weight/bias, linear1.*, linear2.*, norm1.*, norm2.*
plus head.weight/bias (fc_mean or v_out) and, for the actor, log_std.
"""
import math

import torch

TOKENS = ("state_self", "state_others", "cylinders")


def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def encoder(p, obs):
    """obs: dict of [R, n_tokens_k, dim_k] -> features [R, 128]."""
    toks = [obs[k] @ p[f"split_embed.embed.{k}.weight"].T + p[f"split_embed.embed.{k}.bias"] for k in TOKENS]
    x = torch.cat(toks, dim=-2)                                                       # [R, N, 128]
    x = _ln(x, p["split_embed.layer_norm.weight"], p["split_embed.layer_norm.bias"])
    d = x.shape[-1]
    wq, wk, wv = p["attn.in_proj_weight"].split(d, 0)
    bq, bk, bv = p["attn.in_proj_bias"].split(d, 0)
    x0 = x[:, 0]
    q = x0 @ wq.T + bq                                                                # [R, 128]
    k = x @ wk.T + bk                                                                 # [R, N, 128]
    v = x @ wv.T + bv
    s = torch.einsum("rd,rnd->rn", q, k) / math.sqrt(d)
    a = torch.softmax(s, dim=-1)
    o = torch.einsum("rn,rnd->rd", a, v)
    attn = o @ p["attn.out_proj.weight"].T + p["attn.out_proj.bias"]
    y = _ln(x0 + attn, p["norm1.weight"], p["norm1.bias"])                            # norm_first=False
    h = y @ p["linear1.weight"].T + p["linear1.bias"]
    h = 0.5 * h * (1.0 + torch.erf(h / math.sqrt(2.0)))                               # F.gelu (exact)
    y2 = _ln(y + h @ p["linear2.weight"].T + p["linear2.bias"], p["norm2.weight"], p["norm2.bias"])
    return y2                                                                          # mean over the 1-token query dim


def head(p, feat):
    return feat @ p["head.weight"].T + p["head.bias"]


def actor(p, obs, eps=None):
    """-> (action, log_prob [R,1], mean).  eps None = deterministic (mode)."""
    mean = head(p, encoder(p, obs))
    std = torch.exp(p["log_std"]).expand_as(mean)
    action = mean if eps is None else mean + std * eps
    logp = (-((action - mean) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1, keepdim=True)
    return action, logp, mean


def critic(p, obs):
    return head(p, encoder(p, obs))


def philox_normal(seed: int, step: int, num_rows: int, head_dim: int):
    """The noise hs_policy_forward draws itself (csrc/hs_policy.cuh): Philox4x32-10 (oracle/reset_sampler.py, pinned
    by the Random123 known-answer vectors), key = seed, counter = (row, step, block of four head columns);
    Box-Muller on u0 in (0,1], u1 in [0,1).  fp32 like the kernel; log/sin/cos differ from CUDA's by ulps."""
    import numpy as np
    from oracle.reset_sampler import philox4x32_10
    rows = np.arange(num_rows, dtype=np.uint64)
    out = np.zeros((num_rows, 4 * ((head_dim + 3) // 4)), np.float32)
    for blk in range((head_dim + 3) // 4):
        ctr = np.stack([(rows & np.uint64(0xFFFFFFFF)), rows >> np.uint64(32),
                        np.full(num_rows, step & 0xFFFFFFFF, np.uint64),
                        np.full(num_rows, (((step >> 32) << 1) | blk) & 0xFFFFFFFF, np.uint64)], -1).astype(np.uint32)
        key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32), (num_rows, 2))
        u = philox4x32_10(ctr, key)
        f = np.float32
        s24 = f(2.0 ** -24)
        u0 = ((u[:, 0] >> np.uint32(8)).astype(f) + f(1)) * s24
        u1 = (u[:, 1] >> np.uint32(8)).astype(f) * s24
        u2 = ((u[:, 2] >> np.uint32(8)).astype(f) + f(1)) * s24
        u3 = (u[:, 3] >> np.uint32(8)).astype(f) * s24
        r0, r1 = np.sqrt(f(-2) * np.log(u0)).astype(f), np.sqrt(f(-2) * np.log(u2)).astype(f)
        t0, t1 = (f(6.283185307179586) * u1).astype(f), (f(6.283185307179586) * u3).astype(f)
        out[:, 4 * blk + 0], out[:, 4 * blk + 1] = r0 * np.cos(t0), r0 * np.sin(t0)
        out[:, 4 * blk + 2], out[:, 4 * blk + 3] = r1 * np.cos(t1), r1 * np.sin(t1)
    return torch.from_numpy(out[:, :head_dim].copy())
