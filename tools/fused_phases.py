#!/usr/bin/env python
"""Phase timeline of hs_tick_tp_fused_kernel (CTA 0, %globaltimer at the block barriers).  Needs a debug build:
    python multi-uav-pursuit-evasion_b200/build.py --force --define HS_FUSED_TIMING   (then rebuild without it)"""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402
from mupe_b200._lib import lib  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
    raw = C.CDLL(lib._name)
    cfg = mupe_b200.build_hs_config(E)
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    eng = mupe_b200.HsEngine(cfg, dev)
    a = 0.9 / 2 ** 0.5
    dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
    tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
    rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
    cyl = torch.zeros(E, 5, 3, device=dev); cyl[..., 2] = -20.0
    eng.reset(None, dpos, rot, tpos, cyl)
    w = eng.tp_weights(tp)
    eng.step_post_tp(w)
    act = torch.randn(E, 3, 4, device=dev)
    for _ in range(20):
        eng.step_fused(act, w)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 32)()
    assert raw.hs_debug_times(buf) == 0
    t = list(buf)
    names = {1: "tick body / weights g2s done (thread 0 = tick warp 0)", 2: "after block barrier", 3: "weights smem -> TMEM",
             4: "x staged + barrier", 21: "FC + tanh + barrier", 22: "row assembly (thread 0)", 20: "row stores"}
    names.update({5 + s: f"LSTM step {s}" for s in range(10)})
    prev = t[0]
    for i in [1, 2, 3, 4] + list(range(5, 15)) + [21, 22, 20]:
        print(f"{names[i]:55s} +{(t[i] - prev) / 1e3:7.2f} us   (t = {(t[i] - t[0]) / 1e3:7.2f})")
        prev = t[i]
    eng.close()


if __name__ == "__main__":
    main()
