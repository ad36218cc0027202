#!/usr/bin/env python
"""Phase timeline of one tick (t = 8) inside hs_rollout_fused_kernel - the one-tick-per-pass variant - (CTA 0, %globaltimer).  Needs the timing build:
    python multi-uav-pursuit-evasion_b200/build.py --define HS_FUSED_TIMING --out multi-uav-pursuit-evasion_b200/libhs_b200_timing.so
    HS_B200_LIB=multi-uav-pursuit-evasion_b200/libhs_b200_timing.so python tools/rollout_phases.py"""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402
from mupe_b200._lib import lib  # noqa: E402


def main():
    E, T = 4096, 64
    dev = torch.device("cuda:0")
    raw = C.CDLL(lib._name)
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev, rollout_steps=T)
    eng.set_rollout_variant(1)          # the stamps live in hs_rollout_fused_kernel (one tick per predictor pass)
    a = 0.9 / 2 ** 0.5
    dpos = torch.rand(E, 3, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([0.1, -a + 0.1, 0.5], device=dev)
    tpos = torch.rand(E, 3, device=dev) * torch.tensor([a - 0.2, 2 * a - 0.2, 0.2], device=dev) + torch.tensor([-a + 0.1, -a + 0.1, 0.5], device=dev)
    rot = torch.zeros(E, 3, 4, device=dev); rot[..., 0] = 1
    cyl = torch.zeros(E, 5, 3, device=dev); cyl[..., 2] = -20.0
    eng.reset(None, dpos, rot, tpos, cyl)
    w = eng.tp_weights(tp)
    eng.step_post_tp(w)
    acts = torch.randn(T, E, 3, 4, device=dev)
    for _ in range(3):
        eng.rollout_fused(acts, T, w)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 32)()
    assert raw.hs_debug_times(buf) == 0
    t = list(buf)
    us = lambda i, j: (t[i] - t[j]) / 1e3
    print("predictor warps, tick 8 (thread 0):")
    print(f"  wait for the tick warps ('tick done')   {us(1, 0):7.2f} us")
    print(f"  state load + x staging + barrier        {us(2, 1):7.2f} us")
    print(f"  recurrence (10 LSTM steps) + barrier    {us(3, 2):7.2f} us")
    print(f"  FC + tanh + rows + stores + barrier     {us(4, 3):7.2f} us")
    print(f"  whole tick period (loop top to loop top){us(6, 0):7.2f} us")
    print("tick warps, tick 8 (first lane of tick warp 0):")
    print(f"  wait for 'tile free'                    {us(11, 10):7.2f} us")
    print(f"  table + hs_tick_body                    {us(12, 11):7.2f} us")
    print(f"  tick warps start their wait {us(10, 0):+7.2f} us relative to the predictor warps' loop top")
    # one LSTM step inside the recurrence (tick 8): the issuer's wait on h of step 3 -> the MMAs of step 4 -> epilogue of step 4
    for hf in (0, 1):
        b = 16 + 4 * hf
        print(f"half {hf}: issue of step 4 (30 MMAs per M-tile + commit) {us(b + 1, b):6.2f} us | accumulator ready for the epilogue "
              f"{us(b + 2, b + 1):+6.2f} us after the commit | epilogue (tcgen05.ld, gates, h, fences) {us(b + 3, b + 2):6.2f} us")
    print(f"half 0: h of step 4 seen by the issuer {us(24, 19):+6.2f} us after the epilogue's last fence; step period (h3 -> h4 seen) {us(24, 16):6.2f} us")
    print(f"half 1: step period {us(28, 20):6.2f} us; half 1 lags half 0 by {us(20, 16):6.2f} us")
    eng.close()


if __name__ == "__main__":
    main()
