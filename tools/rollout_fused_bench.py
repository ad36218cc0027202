#!/usr/bin/env python
"""hs_rollout_fused (T ticks in one launch) against one hs_step_fused launch per tick (graph of T kernel nodes), same
batches, rollout-storage engines rotating over R independent batches.  Usage: python tools/rollout_fused_bench.py [E] [T] [R]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402
from mupe_b200.engine import RotatingRolloutGraph  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    R = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    dev = torch.device("cuda:0")
    cfg = mupe_b200.build_hs_config(E)
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    a = 0.9 / 2 ** 0.5
    engs = []
    for r in range(R):
        eng = mupe_b200.HsEngine(cfg, dev, rollout_steps=T)
        dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
        tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
        rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
        cyl = torch.zeros(E, 5, 3, device=dev); cyl[..., 0] = torch.arange(5, device=dev) * 0.2; cyl[..., 2] = -20.0
        eng.reset(None, dpos, rot, tpos, cyl)
        eng.step_post_tp(eng.tp_weights(tp))
        engs.append(eng)
    acts = torch.randn(T, E, 3, 4, device=dev)
    ws = [e.tp_weights(tp) for e in engs]

    def timed(fn, reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record(); torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps

    out = {"E": E, "T": T, "rotating_batches": R}
    for v in (1, 2, 3):
        for e in engs:
            e.set_rollout_variant(v)
        us = timed(lambda i: engs[i % R].rollout_fused(acts, T, ws[i % R]), 4 * R)
        out[f"rollout_fused_{v}_ticks_per_pass"] = {"us_per_tick": us / T, "env_steps_per_s": E * T / (us * 1e-6)}
    for e in engs:
        e.set_rollout_variant(0)
    us = timed(lambda i: engs[i % R].rollout_fused(acts, T, ws[i % R]), 4 * R)
    out["rollout_fused"] = {"us_per_rollout": us, "us_per_tick": us / T, "env_steps_per_s": E * T / (us * 1e-6)}
    us = timed(lambda i: engs[i % R].rollout_fused(acts[0], T, ws[i % R]), 4 * R)
    out["rollout_fused_constant_action"] = {"us_per_rollout": us, "us_per_tick": us / T, "env_steps_per_s": E * T / (us * 1e-6)}

    def per_tick(i):
        e, w = engs[i % R], ws[i % R]
        for t in range(T):
            e.step_fused(acts[t], w)
    g = torch.cuda.CUDAGraph()
    # one graph per engine of T hs_step_fused launches
    graphs = []
    for r in range(R):
        e, w = engs[r], ws[r]
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=torch.cuda.Stream(dev)):
            for t in range(T):
                e.step_fused(acts[t], w)
        graphs.append(gr)
    us = timed(lambda i: graphs[i % R].replay(), 4 * R)
    out["graph_of_T_step_fused"] = {"us_per_rollout": us, "us_per_tick": us / T, "env_steps_per_s": E * T / (us * 1e-6)}
    print(json.dumps(out), flush=True)
    for e in engs:
        e.close()


if __name__ == "__main__":
    main()
