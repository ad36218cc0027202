#!/usr/bin/env python
"""SURVEY 8f row 3 measurement: hs_policy_forward (one launch per network) against the same network written with
torch.nn modules the way the reference builds it (SplitEmbedding + nn.MultiheadAttention + feed-forward block + head,
eager fp32 on the same GPU, TF32 off).  Device time per call, rows/s, fp32 FLOP/s of the fused kernel's arithmetic.
Usage: python tools/policy_bench.py [E ...]"""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402


class EagerEncoder(nn.Module):                       # networks.py:126-161, 249-314 with stock torch modules
    def __init__(self, D, head_dim):
        super().__init__()
        self.e_self, self.e_others, self.e_cyl = nn.Linear(D, 128), nn.Linear(3, 128), nn.Linear(5, 128)
        self.ln = nn.LayerNorm(128)
        self.attn = nn.MultiheadAttention(128, 1, batch_first=True)
        self.linear1, self.linear2 = nn.Linear(128, 128), nn.Linear(128, 128)
        self.norm1, self.norm2 = nn.LayerNorm(128), nn.LayerNorm(128)
        self.head = nn.Linear(128, head_dim)
        self.log_std = nn.Parameter(torch.zeros(head_dim))

    def forward(self, s, o, c, eps):
        x = self.ln(torch.cat([self.e_self(s), self.e_others(o), self.e_cyl(c)], dim=-2))
        shp = x.shape[:-2]
        x = x.reshape(-1, x.shape[-2], x.shape[-1])
        y = self.norm1(x[:, [0]] + self.attn(x[:, [0]], x, x, need_weights=False)[0])
        y = self.norm2(y + self.linear2(F.gelu(self.linear1(y))))
        feat = y.mean(-2).reshape(*shp, -1)
        mean = self.head(feat)
        dist = torch.distributions.Independent(torch.distributions.Normal(mean, self.log_std.exp().expand_as(mean)), 1)
        a = mean + self.log_std.exp() * eps
        return a, dist.log_prob(a).unsqueeze(-1)

    def params(self):
        sd = self.state_dict()
        m = {"e_self": "split_embed.embed.state_self", "e_others": "split_embed.embed.state_others",
             "e_cyl": "split_embed.embed.cylinders", "ln": "split_embed.layer_norm"}
        out = {}
        for k, v in sd.items():
            head, _, tail = k.partition(".")
            out[(m[head] + "." + tail) if head in m else k] = v
        return out


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    Es = [int(x) for x in sys.argv[1:]] or [4096, 16384, 65536]
    A, D = 3, 35
    torch.manual_seed(0)
    ref = EagerEncoder(D, 4).to(dev)
    net = mupe_b200.FusedPolicy(ref.params(), 2, 3, dev)
    mac_per_row = D * 128 + 4 * 128 * 128 + 5 * (4 * 128) + 7 * 128 * 2 + 4 * 128   # fused kernel's arithmetic
    for E in Es:
        s, o, c = torch.randn(E, A, 1, D, device=dev), torch.randn(E, A, 2, 3, device=dev), torch.randn(E, A, 3, 5, device=dev)
        eps = torch.randn(E, A, 4, device=dev)
        out = {}
        us_ffma = timed(lambda: net(s, o, c, eps=eps, out=out, impl=1))
        us = timed(lambda: net(s, o, c, eps=eps, out=out, impl=2))
        with torch.no_grad():
            us_ref = timed(lambda: ref(s, o, c, eps), reps=5)
            a_ref, lp_ref = ref(s, o, c, eps)
        err = (out["action"] - a_ref).abs().max().item()
        R = E * A
        print(json.dumps({"what": "hs_policy_forward", "E": E, "rows": R, "us_tcgen05": us, "us_ffma": us_ffma,
                          "rows_per_s": R / (us * 1e-6),
                          "fp32_equiv_TFLOPs": 2 * mac_per_row * R / (us * 1e-6) / 1e12,
                          "tf32_TFLOPs_executed": 3 * 2 * (40 * 128 + 4 * 128 * 128) * R / (us * 1e-6) / 1e12,
                          "ffma_TFLOPs": 2 * mac_per_row * R / (us_ffma * 1e-6) / 1e12, "torch_eager_us": us_ref,
                          "speedup": us_ref / us, "max_abs_action_diff_vs_eager": err}), flush=True)


if __name__ == "__main__":
    main()
