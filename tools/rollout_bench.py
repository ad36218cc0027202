#!/usr/bin/env python
"""SURVEY 8f row 4 measurements: (a) hs_gae against the reference's eager loop on the same GPU (device time,
algorithmic GB/s = 5 fp32 arrays + done per element); (b) the collector with the time-major RolloutStorage against the
generic clone-per-step + stack collector (env-steps/s through TransformedEnv.step with a random policy).
Usage: python tools/rollout_bench.py [E] [T]"""
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mupe_b200  # noqa: E402
from mupe_b200 import rollout as R  # noqa: E402


def ref_gae(reward, done, value, next_value, gamma, lmbda):      # the reference's loop, gae.py:27-51
    not_done = 1.0 - done.float()
    T = not_done.shape[1]
    gae, adv = 0, torch.zeros_like(reward)
    for step in reversed(range(T)):
        delta = reward[:, step] + gamma * next_value * not_done[:, step] - value[:, step]
        adv[:, step] = gae = delta + (gamma * lmbda * not_done[:, step] * gae)
        next_value = value[:, step]
    return adv, adv + value


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    A = 3
    dev = torch.device("cuda:0")
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(0)
    for layout in ("time_major", "env_major"):
        shp = (T, E, A, 1) if layout == "time_major" else (E, T, A, 1)
        mk = lambda: torch.randn(*shp, generator=g, device=dev)
        reward, value = mk(), mk()
        done = torch.rand(*shp[:2], 1, generator=g, device=dev) < 0.01
        if layout == "time_major":
            reward, value, done = reward.transpose(0, 1), value.transpose(0, 1), done.transpose(0, 1)
        nv = torch.randn(E, A, 1, generator=g, device=dev)
        nbytes = E * T * (A * 5 * 4 + 1)
        us = timed(lambda: R.compute_gae(reward, done, value, nv, 0.995, 0.95))
        us_n = timed(lambda: R.compute_gae(reward, done, value, nv, 0.995, 0.95, normalize=True))
        d4 = done.unsqueeze(2).expand(E, T, A, 1)
        us_ref = timed(lambda: ref_gae(reward, d4, value, nv, 0.995, 0.95), reps=3)
        print(json.dumps({"what": "hs_gae", "layout": layout, "E": E, "T": T, "A": A, "us": us, "us_with_normalize": us_n,
                          "algorithmic_GBps": nbytes / (us * 1e-6) / 1e9, "frac_of_measured_hbm": nbytes / (us * 1e-6) / 1e9 / peak,
                          "torch_eager_loop_us": us_ref, "speedup": us_ref / us}), flush=True)

    Ec, Tc = 4096, 64
    for steps in (0, Tc):
        cfg = mupe_b200.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": Ec, "task.sim.device": "cuda:0",
                                                                     "task.env.rollout_steps": steps})
        base = mupe_b200.IsaacEnv.REGISTRY[cfg.task.name.lower()](cfg, headless=True)
        env = mupe_b200.TransformedEnv(base, mupe_b200.Compose(mupe_b200.InitTracker(), mupe_b200.PIDRateController()))
        act = torch.randn(Ec, 3, 4, device=dev)

        def policy(td):
            td.set(("agents", "action"), act)
            return td
        col = mupe_b200.SyncDataCollector(env, policy=policy, frames_per_batch=Ec * Tc, total_frames=-1, return_same_td=True)
        it = iter(col)
        next(it); next(it)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            d = next(it)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"what": "collector", "mode": "rollout_storage" if steps else "clone_and_stack", "E": Ec, "T": Tc,
                          "env_steps_per_s": n * Ec * Tc / dt, "ms_per_rollout": 1e3 * dt / n}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
