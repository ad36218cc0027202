#!/usr/bin/env python
"""Experiment: the host-to-host tick (H2D action, tick, predictor, D2H observation + reward + done) as ONE CUDA graph
launch with memcpy nodes, against hs_step_host_io (7 stream API calls).  Also times the bare pieces.
Usage: python tools/e2e_graph.py [E]"""
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
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
    h_act = torch.randn(E, 3, 4).pin_memory()
    n = 300
    for _ in range(10):
        eng.step_host(h_act, w)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        eng.step_host(h_act, w)
    t_api = (time.perf_counter() - t0) / n

    # one graph per output set: H2D -> tick -> {D2H tick outputs || predictor -> D2H state_self}
    npol = eng.sets[0].policy_words
    mirror = torch.empty(npol, dtype=torch.float32).pin_memory()
    done_h = torch.empty(E, dtype=torch.uint8).pin_memory()
    staging = torch.empty(E, 3, 4, device=dev)
    import ctypes as C
    from mupe_b200._lib import check, lib
    graphs = []
    side = torch.cuda.Stream(dev)
    side2 = torch.cuda.Stream(dev)
    keep = eng.cur
    for i in range(len(eng.sets)):
        s = eng.sets[i]
        n_self = s["state_self"].numel()
        eng._bind(i, (i - 1) % len(eng.sets))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            cur = torch.cuda.current_stream(dev)
            staging.copy_(h_act, non_blocking=True)
            check(lib.hs_step_pre(eng._h, staging.data_ptr(), 1, None, cur.cuda_stream), "pre")
            side2.wait_stream(cur)
            with torch.cuda.stream(side2):
                mirror[n_self:npol].copy_(s.slab[n_self:npol], non_blocking=True)
                done_h.copy_(s["done"].view(torch.uint8).reshape(E), non_blocking=True)
            check(lib.hs_step_post_tp(eng._h, C.byref(w), None, cur.cuda_stream), "post")
            mirror[:n_self].copy_(s.slab[:n_self], non_blocking=True)
            cur.wait_stream(side2)
        graphs.append(g)
    eng._bind(keep, keep)
    st = torch.cuda.current_stream(dev)

    def gtick():
        eng.cur = (eng.cur + 1) % len(eng.sets)
        graphs[eng.cur].replay()
        st.synchronize()
    for _ in range(10):
        gtick()
    t0 = time.perf_counter()
    for _ in range(n):
        gtick()
    t_graph = (time.perf_counter() - t0) / n

    # bare pieces
    def wall(fn, reps=200):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e6 * (time.perf_counter() - t0) / reps
    s = eng.sets[0]
    n_self = s["state_self"].numel()
    p_d2h_self = wall(lambda: (mirror[:n_self].copy_(s.slab[:n_self], non_blocking=True), st.synchronize()))
    p_d2h_all = wall(lambda: (mirror.copy_(s.slab[:npol], non_blocking=True), st.synchronize()))
    p_h2d = wall(lambda: (staging.copy_(h_act, non_blocking=True), st.synchronize()))
    p_sync = wall(lambda: st.synchronize())
    print(json.dumps({"E": E, "host_io_api_us": 1e6 * t_api, "host_io_graph_us": 1e6 * t_graph,
                      "env_steps_per_s_api": E / t_api, "env_steps_per_s_graph": E / t_graph,
                      "d2h_state_self_us": p_d2h_self, "d2h_policy_prefix_us": p_d2h_all, "h2d_action_us": p_h2d,
                      "empty_sync_us": p_sync, "bytes_state_self": n_self * 4, "bytes_policy_prefix": npol * 4}))
    eng.close()


if __name__ == "__main__":
    main()
