#!/usr/bin/env python
"""Phase timeline of hs_policy_forward_tc_kernel (CTA 0, thread 0, %globaltimer).  Needs the debug build:
    python multi-uav-pursuit-evasion_b200/build.py --force --define HS_FUSED_TIMING   (then rebuild without it)"""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402
from mupe_b200._lib import lib  # noqa: E402
from mupe_b200.policy import FusedPolicy, init_params  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
    dev = torch.device("cuda:0")
    raw = C.CDLL(lib._name)
    torch.manual_seed(0)
    net = FusedPolicy(init_params(35, 2, 3, 4, True, dev), 2, 3, dev)
    s, o, c = torch.randn(R, 1, 35, device=dev), torch.randn(R, 2, 3, device=dev), torch.randn(R, 3, 5, device=dev)
    eps = torch.randn(R, 4, device=dev)
    for _ in range(10):
        net(s, o, c, eps=eps, impl=2)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 32)()
    assert raw.hs_debug_times(buf) == 0
    t = list(buf)
    names = ["inputs staged", "L0 gemm (embed, K=40)", "L0 epilogue (LN, stash, A)", "L1 gemm (W_kq)", "q' + attention + A",
             "L2 gemm (W_ov)", "L2 epilogue (residual, LN1)", "L3 gemm (W1)", "L3 epilogue (gelu)", "L4 gemm (W2)",
             "L4 epilogue (LN2, head, sample)"]
    for i in range(1, 11):
        print(f"{names[i]:40s} +{(t[i] - t[i - 1]) / 1e3:7.2f} us   (t = {(t[i] - t[0]) / 1e3:7.2f})")
    sub = ["q' from D + bias", "x0 from the stash", "partial sums (13 x 32 FMA)", "row exchange", "token statistics + scores",
           "softmax + U", "xbar (10 x 32 FMA)", "tf32 split + tcgen05.st"]
    prev = t[3]
    for k, i in enumerate(range(11, 19)):
        print(f"   attention: {sub[k]:32s} +{(t[i] - prev) / 1e3:7.2f} us")
        prev = t[i]


if __name__ == "__main__":
    main()
