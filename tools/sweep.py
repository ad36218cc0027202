#!/usr/bin/env python
"""Kernel-only sweep over the batch size: device time per launch (CUDA events around a CUDA
graph of back-to-back launches), env-steps/s and achieved algorithmic GB/s vs the measured
HBM peak.  Working sets above the 126 MB L2 are naturally cold; smaller ones rotate over
enough independent batches to exceed it.  Usage: python tools/sweep.py [E ...]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import mupe_b200  # noqa: E402
from mupe_b200._lib import check, lib  # noqa: E402
import ctypes  # noqa: E402


def run(E, dev, peak, reps=20):
    C = int(os.environ.get("HS_SWEEP_C", "5"))          # cylinder capacity, all of them active (config 3: 8)
    cfg = mupe_b200.build_hs_config(E, num_cylinders=C)
    per_batch = bench.algorithmic_bytes(C=C)["total"] * E * 1.3
    R = max(2, min(16, int(2 * 126e6 / per_batch) + 1))
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    engs = []
    for r in range(R):
        eng = mupe_b200.HsEngine(cfg, dev)
        a = 0.9 / 2 ** 0.5
        dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
        tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
        rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
        cyl = torch.zeros(E, C, 3, device=dev)
        cyl[..., :2] = (torch.randint(-3, 4, (E, C, 2), device=dev)).float() * 0.2
        cyl[..., 2] = 0.6
        eng.set_predictor_variant(int(os.environ.get("HS_TP_VARIANT", "-1")))
        eng.set_tick_mapping(int(os.environ.get("HS_TICK_MAPPING", "0")))      # 1: 4 lanes per env, 2: one lane per env
        if os.environ.get("HS_TP_RING", "0") == "1":
            eng.set_tp_ring(True)
        eng.reset(None, dpos, rot, tpos, cyl)
        eng.step_post_tp(eng.tp_weights(tp))
        eng.graph_action = torch.randn(E, 3, 4, device=dev)
        engs.append(eng)
    torch.cuda.synchronize()
    n = max(R, 32 if E <= 65536 else 8)
    out = {"E": E, "num_cylinders": C, "rotating_batches": R, "tick_mapping": int(os.environ.get("HS_TICK_MAPPING", "0")),
           "tp_ring": int(os.environ.get("HS_TP_RING", "0"))}
    ab = bench.algorithmic_bytes(C=C)
    for name, fn, nbytes in (
            ("tick", lambda e, st: check(lib.hs_step_pre(e._h, e.graph_action.data_ptr(), 1, None, st), "pre"), ab["tick"]),
            ("tp_fill", lambda e, st: check(lib.hs_step_post_tp(e._h, ctypes.byref(e.tp_weights(tp)), None, st), "post"), ab["fill"]),
            # the whole tick as hs_step_fused launches it: ONE kernel up to one 32-env tile per SM, tick + predictor above
            ("step_fused", lambda e, st: check(lib.hs_step_fused(e._h, e.graph_action.data_ptr(), 1, None,
                                                                 ctypes.byref(e.tp_weights(tp)), None, st), "fused"), ab["total"])):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        with torch.cuda.graph(g, stream=side):
            st = torch.cuda.current_stream(dev).cuda_stream
            for i in range(n):
                fn(engs[i % R], st)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (reps * n)
        gbs = nbytes * E / (us * 1e-6) / 1e9
        out[name] = {"us_per_launch": us, "env_steps_per_s": E / (us * 1e-6), "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak}
    t = out["tick"]["us_per_launch"] + out["tp_fill"]["us_per_launch"]
    out["tick_plus_tp"] = {"us": t, "env_steps_per_s": E / (t * 1e-6), "algorithmic_GBps": ab["total"] * E / (t * 1e-6) / 1e9}
    for e in engs:
        e.close()
    return out


def main():
    dev = torch.device("cuda:0")
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    Es = [int(x) for x in sys.argv[1:]] or [4096, 16384, 65536, 262144, 1048576]
    for E in Es:
        print(json.dumps(run(E, dev, peak)), flush=True)


if __name__ == "__main__":
    main()
