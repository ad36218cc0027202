"""Generates tests/golden/envgen_episodes.npz by running the REFERENCE'S OWN HideAndSeek_envgen source
(omni_drones/envs/hide_and_seek/hideandseek_envgen.py: `_reset_idx` with the particle generator :875-1013, the tick, and
`_compute_reward_and_done` with the archive bookkeeping :1241-1333, plus its `GenBuffer` class :209-377) through
oracle/ref_harness.RefEnvgen for several generator cycles of short episodes.  Test infrastructure; run in the build
container (needs /root/reference or baseline/_ref):  python -m oracle.gen_envgen_episode_golden

Recorded per reset: the task table the reference drew (`all_tasks`), `num_unif`, the sampled orientations; per tick: the
raw action, observation / reward / done and every stats key; per episode end: the archive (`_history_buffer`), the
averaged weights, `ratio_unif`, `update_iter`.  A replay injects the recorded draws (uniform tasks, perturbed archive tasks,
orientations, actions) into the implementation under test and must reproduce everything else."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import hs_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

CFG = dict(E=24, L=5, eval_iter=2, cycles=5, ratio_unif=0.3, R_min=0.5, R_max=1.0, success_threshold=0.6,
           expand_cylinders=True, expand_step=0.05, buffer_length=20, catch_radius=0.55, min_cylinders=4, seed=3)


def main(out_path=None):
    c = dict(CFG)
    for kv in os.environ.get("ENVGEN_GOLDEN_OVERRIDE", "").split(","):        # tuning aid: k=v,k=v
        if "=" in kv:
            k, v = kv.split("=")
            c[k] = type(CFG[k])(float(v))
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    P = O.HSParams(max_episode_length=c["L"], catch_radius=c["catch_radius"])
    tp_ref = RH.load_reference()["TP_net"](input_dim=P.tp_frame_dim, output_dim=3 * P.future_step,
                                           future_predcition_step=P.future_step, window_step=1)
    tp_sd = {k: v.clone() for k, v in tp_ref.state_dict().items()}
    r = RH.RefEnvgen(P, c["E"], eval_iter=c["eval_iter"], ratio_unif=c["ratio_unif"], R_min=c["R_min"], R_max=c["R_max"],
                     success_threshold=c["success_threshold"], expand_cylinders=c["expand_cylinders"],
                     expand_step=c["expand_step"], buffer_length=c["buffer_length"], min_cylinders=c["min_cylinders"],
                     tp_state_dict=tp_sd)
    e = r.env
    E, A = c["E"], P.num_agents
    out = {f"cfg_{k}": np.asarray(v) for k, v in c.items()}
    out.update({f"tp_{k}": v.numpy() for k, v in tp_sd.items()})
    out["stat_keys"] = np.array(r.stat_keys)
    g = torch.Generator().manual_seed(17)
    ep = 0
    for cyc in range(c["cycles"]):
        for it in range(c["eval_iter"]):
            td = r.reset_all()
            pre = f"ep{ep}_"
            out[pre + "all_tasks"] = np.asarray(e.all_tasks, dtype=np.float64)
            out[pre + "num_unif"] = np.asarray(e.num_unif)
            out[pre + "drone_rot"] = r.store["drot"].numpy().copy()
            out[pre + "reset_state_self"] = td[("agents", "observation", "state_self")].numpy().copy()
            out[pre + "active_cylinders"] = e.active_cylinders.numpy().copy()
            done = torch.zeros(E, 1, dtype=torch.bool)
            for t in range(c["L"]):
                act = torch.randn(E, A, 4, generator=g) * 0.7
                nxt, _ = r.step(act, done)
                done = nxt["done"].clone()
                tp = pre + f"t{t}_"
                out[tp + "action"] = act.numpy()
                out[tp + "state_self"] = nxt[("agents", "observation", "state_self")].numpy().copy()
                out[tp + "reward"] = nxt[("agents", "reward")].numpy().copy()
                out[tp + "done"] = done.numpy().copy()
                out[tp + "stats"] = np.stack([nxt["stats"][k].reshape(E).float().numpy() for k in r.stat_keys])
            assert bool(done.all())
            out[pre + "history"] = np.asarray(e.gen_buffer._history_buffer, dtype=np.float64).copy()
            out[pre + "weights"] = np.asarray(e.gen_buffer._weight_buffer, dtype=np.float64).reshape(-1).copy()
            out[pre + "ratio_unif"] = np.asarray(float(e.ratio_unif))
            out[pre + "update_iter"] = np.asarray(int(e.update_iter))
            print(f"episode {ep}: num_unif {int(e.num_unif)}, success {float(e.stats['success'].mean()):.3f}, "
                  f"history {len(e.gen_buffer._history_buffer)}, ratio_unif {e.ratio_unif}, update_iter {e.update_iter}")
            ep += 1
    out["num_episodes"] = np.asarray(ep)
    out_path = out_path or os.path.join(REPO, "tests", "golden", "envgen_episodes.npz")
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
