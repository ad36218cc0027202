#!/usr/bin/env python
"""Launches the tick (+ predictor) a few times at a given batch size without CUDA graphs, as a
target for `ncu -k regex:hs_tick --set full --import-source on`.  Usage: python tools/ncu_target.py E [reps] [C]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    cfg = mupe_b200.build_hs_config(E, num_cylinders=C)
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    eng = mupe_b200.HsEngine(cfg, dev)
    eng.set_tick_mapping(int(os.environ.get("HS_TICK_MAPPING", "0")))      # 1: 4 lanes per env, 2: one lane per env
    if os.environ.get("HS_TP_RING", "0") == "1":
        eng.set_tp_ring(True)                                              # TP window as a ring (hs_buffers.tp_ring)
    a = 0.9 / 2 ** 0.5
    dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
    tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
    rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
    cyl = torch.zeros(E, C, 3, device=dev)
    cyl[..., :2] = (torch.randint(-3, 4, (E, C, 2), device=dev)).float() * 0.2
    cyl[..., 2] = 0.6
    eng.reset(None, dpos, rot, tpos, cyl)
    w = eng.tp_weights(tp)
    eng.step_post_tp(w)
    act = torch.randn(E, 3, 4, device=dev)
    for _ in range(reps):
        eng.step_pre(act, raw=True, reset_pid=None)
        eng.step_post_tp(w)
    torch.cuda.synchronize()
    eng.close()


if __name__ == "__main__":
    main()
