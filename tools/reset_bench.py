"""Episode-boundary cost: the reset pose sampler as one CUDA kernel
(hs_sample_reset) against the vectorised torch sampler it replaces, and the whole env.reset().
Usage: python tools/reset_bench.py [E ...]   -> one JSON line per batch size."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mupe_b200 as m


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 65536]
    for E in sizes:
        row = {"E": E}
        for name, flag in (("device_kernel", 1), ("torch_sampler", 0)):
            cfg = m.compose("HideAndSeek", "mappo", overrides={
                "task.env.num_envs": E, "task.use_random_cylinder": 1, "task.cylinder.max_num": 8,
                "task.env.device_reset_sampler": flag, "algo.use_TP_net": 0})
            env = m.IsaacEnv.REGISTRY[cfg.task.name](cfg, headless=True)
            reps = 20
            row[name] = {"sample_us": timed(lambda: env._sample_reset(E), reps),
                         "env_reset_us": timed(lambda: env.reset(), reps)}
            env.close()
        row["sample_speedup"] = row["torch_sampler"]["sample_us"] / row["device_kernel"]["sample_us"]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
