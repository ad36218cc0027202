#!/usr/bin/env python
"""Two hs_rollout_fused launches (T ticks each, rollout-storage engine) as a target for
`ncu -k regex:hs_rollout_fused --set full`.  Usage: python tools/ncu_target_rollout.py [E] [T]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev, rollout_steps=T)
    a = 0.9 / 2 ** 0.5
    dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
    tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
    rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
    cyl = torch.zeros(E, 5, 3, device=dev); cyl[..., 0] = torch.arange(5, device=dev) * 0.2; cyl[..., 2] = -20.0
    eng.reset(None, dpos, rot, tpos, cyl)
    w = eng.tp_weights(tp)
    eng.step_post_tp(w)
    acts = torch.randn(T, E, 3, 4, device=dev)
    for _ in range(2):
        eng.rollout_fused(acts, T, w)
    torch.cuda.synchronize()
    eng.close()


if __name__ == "__main__":
    main()
