"""Generates tests/golden/hover.npz by executing the REFERENCE'S OWN Hover source (build container only):
omni_drones/envs/single/hover.py `_reset_idx` / `_pre_sim_step` / `_compute_state_and_obs` / `_compute_reward_and_done`
through oracle/ref_harness.RefHover.  Every recorded tick stores the complete pre-tick state (vehicle + the task's running
values), the action and every output, so that consumers replay single ticks.  oracle/hover_oracle.py is checked against
the reference tick by tick while generating.

    python -m oracle.gen_hover_golden
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import hover_oracle as HO      # noqa: E402
from oracle import hs_oracle as O          # noqa: E402
from oracle.gen_golden import check        # noqa: E402
from oracle.ref_harness import RefHover    # noqa: E402


def task_state(h):
    e = h.env
    return torch.cat([e.last_linear_v, e.last_angular_v, e.last_linear_a, e.last_angular_a, e.last_linear_jerk, e.last_angular_jerk,
                      e.linear_v_episode, e.angular_v_episode, e.linear_a_episode, e.angular_a_episode, e.linear_jerk_episode,
                      e.angular_jerk_episode], dim=-1).detach().clone().float()


def vehicle_state(h):
    s, d, t = h.base.store, h.base.drone, h.base.transform
    E = h.E
    ctl = t.controller
    return dict(pos=s["dpos"], quat=s["drot"], linvel=s["dvel"][..., :3], angvel=s["dvel"][..., 3:], throttle=d.throttle,
                integ=getattr(ctl, "integ", torch.zeros(E, 3)).reshape(E, 1, 3),
                last_rate=getattr(ctl, "last_body_rate", torch.zeros(E, 3)).reshape(E, 1, 3),
                prev_action=h.env.info["prev_action"], progress=h.env.progress_buf)


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("needs the reference tree at /root/reference (build container only)")
    torch.manual_seed(0)
    E, ticks = 12, 6
    g = torch.Generator().manual_seed(7)
    h = RefHover(E)
    pos = torch.tensor([-1.0, -1.0, 0.05]) + torch.tensor([2.0, 2.0, 1.95]) * torch.rand(E, 1, 3, generator=g)
    pos[0, 0] = torch.tensor([0.0, 0.0, 1.005])                    # one env inside the 2 cm position bonus
    rpy = torch.tensor([-0.2, -0.2, 0.0]) * torch.pi + torch.tensor([0.4, 0.4, 0.5]) * torch.pi * torch.rand(E, 1, 3, generator=g)
    rpy[0] = 0.0
    rot = O.euler_to_quat(rpy)
    rec = {"init/pos": pos.numpy(), "init/rot": rot.numpy()}
    obs = h.reset_with(torch.ones(E, dtype=torch.bool), pos, rot)
    rec["reset/obs"] = obs["agents"]["observation"].detach().float().numpy()
    rec["reset/stats"] = h.stats_matrix().numpy()
    rec["reset/task_state"] = task_state(h).numpy()
    rec["reset/target_heading"] = h.env.target_heading.detach().clone().numpy()
    st0, ts0 = torch.zeros(E, 39), torch.zeros(E, 12)
    o0, _, _ = HO.hover_post(torch.cat([pos, rot, torch.zeros(E, 1, 6)], -1).reshape(E, 13), torch.zeros(E), st0, ts0,
                             h.env.target_heading.reshape(E, 3), with_reward=False)
    check("hover/reset/obs", obs["agents"]["observation"].reshape(E, -1), o0)
    check("hover/reset/stats", h.stats_matrix(), st0)
    h.env.progress_buf[:] = 496.0                                  # the done tick (progress 500) falls inside the run
    done_prev = torch.zeros(E, dtype=torch.bool)
    for t in range(ticks):
        act = torch.randn(E, 1, 4, generator=g) * (1.2 if t % 2 == 0 else 0.3)
        pre = {k: v.detach().clone().float() for k, v in vehicle_state(h).items()}
        pre_stats, pre_task = h.stats_matrix(), task_state(h)
        obs, rd, aux = h.step(act, done_prev)
        post = {k: v.detach().clone().float() for k, v in vehicle_state(h).items()}
        out = dict(obs=obs["agents"]["observation"], reward=rd["agents"]["reward"], done=rd["done"].float(),
                   stats=h.stats_matrix(), task_state=task_state(h), drone_state=obs["info"]["drone_state"],
                   throttle_diff=h.base.drone.throttle_difference, cmds=aux["cmds"], ctbr=aux["ctbr"], target_rate=aux["target_rate"])
        out = {k: v.detach().clone().float() for k, v in out.items()}
        # the restatement on the reference's post-tick drone state
        st, ts = pre_stats.clone(), pre_task.clone()
        o, r, d = HO.hover_post(out["drone_state"].reshape(E, 13), post["progress"], st, ts, h.env.target_heading.reshape(E, 3),
                                out["cmds"].reshape(E, 4), out["ctbr"].reshape(E, 4), out["target_rate"].reshape(E, 3),
                                out["throttle_diff"].reshape(E))
        check(f"hover/t{t}/obs", out["obs"].reshape(E, -1), o)
        check(f"hover/t{t}/reward", out["reward"].reshape(E), r)
        check(f"hover/t{t}/stats", out["stats"], st, rtol=2e-5, atol=2e-5)
        check(f"hover/t{t}/task_state", out["task_state"], ts, rtol=2e-5, atol=2e-4)
        assert torch.equal(out["done"].reshape(E).bool(), d)
        rec[f"t{t}/action"], rec[f"t{t}/done_prev"] = act.numpy(), done_prev.numpy()
        for k, v in pre.items():
            rec[f"t{t}/pre/{k}"] = v.numpy()
        rec[f"t{t}/pre/stats"], rec[f"t{t}/pre/task_state"] = pre_stats.numpy(), pre_task.numpy()
        for k, v in out.items():
            rec[f"t{t}/out/{k}"] = v.numpy()
        for k, v in post.items():
            rec[f"t{t}/post/{k}"] = v.numpy()
        done_prev = rd["done"].reshape(-1).clone()
    rec["meta/E"], rec["meta/ticks"] = np.array(E), np.array(ticks)
    path = os.path.join(REPO, "tests", "golden", "hover.npz")
    np.savez_compressed(path, **rec)
    print(f"hover: reference == restatement on reset + {ticks} ticks (done tick included); wrote {path} ({os.path.getsize(path) // 1024} KiB)")


if __name__ == "__main__":
    main()
