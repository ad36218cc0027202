"""Host-side profile of the end-to-end env.step() loop (where the time between device ticks goes)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mupe_b200 as m


def main():
    E = 4096
    cfg = m.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": E, "algo.use_TP_net": 1})
    base = m.IsaacEnv.REGISTRY[cfg.task.name](cfg, headless=True)
    env = m.TransformedEnv(base, m.Compose(m.InitTracker(), m.PIDRateController()))
    td = env.reset()
    h_act = torch.randn(E, 3, 4).pin_memory()
    d_act = torch.empty(E, 3, 4, device=base.device)
    out = base.engine.out
    h_res = torch.empty(out.policy_words, dtype=torch.float32).pin_memory()

    def step(td):
        d_act.copy_(h_act, non_blocking=True)
        td.set(("agents", "action"), d_act)
        td = env.step(td)
        o = base.engine.out
        h_res.copy_(o.slab[:o.policy_words], non_blocking=True)
        torch.cuda.synchronize()
        return m.step_mdp(td)
    for _ in range(20):
        td = step(td)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        td = step(td)
    pr.disable()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
