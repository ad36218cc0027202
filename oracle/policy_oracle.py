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
