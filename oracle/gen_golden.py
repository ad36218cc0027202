"""Generates tests/golden/*.npz by executing the REFERENCE'S OWN SOURCE (build container only).

    python -m oracle.gen_golden          # needs /root/reference, writes tests/golden/

For each case the reference code (oracle/ref_harness.py: AST-extracted methods of
hideandseek.py / multirotor.py / transforms.py / lee_position_controller.py / rotor_group.py
run on CPU, with hs_oracle.rigid_body_step standing in for PhysX) is reset from an injected
initial state and stepped with seeded actions.  At every recorded tick we store the complete
pre-tick state, the inputs and every output tensor, so that the consumers (the CPU test of
hs_oracle.py and the GPU test of the CUDA kernels) can replay single ticks from the stored
state and compare with what the reference produced.  While generating, hs_oracle.py is checked
against the reference tick by tick (this is what "pins" the oracle).
"""
import os
import zlib
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import hs_oracle as O            # noqa: E402
from oracle.ref_harness import RefEnv        # noqa: E402

OUT_DIR = os.path.join(REPO, "tests", "golden")

CASES = {
    # name: (params kwargs, E, scenario, min_cylinders, ticks, progress0)
    "default_tp": (dict(), 24, "random_cylinders", 4, 6, None),
    "empty_notp": (dict(use_tp_net=False), 16, "empty", 4, 5, None),
    "wall_tp": (dict(num_cylinders=5), 16, "wall", 4, 5, None),
    "c8_tp": (dict(num_cylinders=8), 16, "random_cylinders", 8, 4, None),
    "done_tick": (dict(), 8, "random_cylinders", 4, 4, 797.0),
    "passage_tp": (dict(num_cylinders=6), 16, "passage", 4, 4, None),
    "narrow_gap_tp": (dict(num_cylinders=5), 16, "narrow_gap", 4, 4, None),
    "deploy_smooth_tp": (dict(use_deployment=True, smoothness_coef=2.0), 12, "random_cylinders", 4, 4, None),   # smoothness reward paid
    "gated_smooth_tp": (dict(use_deployment=False, smoothness_coef=2.0), 12, "random_cylinders", 4, 3, None),   # ... and gated off (hideandseek.py:992-994)
    "a2_tp": (dict(num_agents=2), 12, "random_cylinders", 4, 4, None),
    # smoothness curriculum: base_env.update_epoch is rewritten between ticks (scripts/train_deploy.py:270) and the
    # coefficient min(max, init + smooth_lr * update_epoch) is recomputed at every reward call (hideandseek.py:988-991)
    "deploy_epoch_tp": (dict(use_deployment=True, smoothness_coef=0.5, smooth_lr=0.4, max_smoothness_coef=5.0), 12,
                        "random_cylinders", 4, 4, None),
}
CASES["obstacles_tp"] = (dict(use_obstacles=True), 12, "random_cylinders", 4, 4, None)     # cylinders inside the TP frame
# four pursuers: the reference's fixed scenarios carry four start rows (hideandseek.py:633-682); 19-float TP frame
CASES["a4_narrow_gap_tp"] = (dict(num_agents=4, num_cylinders=5), 12, "narrow_gap", 4, 4, None)
CASES["random6_tp"] = (dict(num_cylinders=6), 12, "random", 4, 4, None)      # the fifth fixed scenario (hideandseek.py:512-521, 663-672)
CASES["a4_random_tp"] = (dict(num_agents=4, num_cylinders=5), 12, "random_cylinders", 4, 4, None)
UPDATE_EPOCHS = {"deploy_epoch_tp": [0, 3, 3, 20]}        # -> 0.5, 1.7, 1.7, min(5, 8.5)


def snapshot_state(ref: RefEnv):
    s, d, e, t = ref.store, ref.drone, ref.env, ref.transform
    ctl = t.controller
    E, A = ref.E, ref.P.num_agents
    integ = getattr(ctl, "integ", torch.zeros(E * A, 3)).reshape(E, A, 3)
    last = getattr(ctl, "last_body_rate", torch.zeros(E * A, 3)).reshape(E, A, 3)
    out = dict(pos=s["dpos"], quat=s["drot"], linvel=s["dvel"][..., :3], angvel=s["dvel"][..., 3:],
               throttle=d.throttle, integ=integ, last_rate=last, prev_action=e.info["prev_action"],
               tpos=s["tpos"][:, 0], tvel=s["tvel"][:, 0, :3], cyl=s["cpos"], progress=e.progress_buf,
               stats=torch.cat([e.stats[k] for k in O.STAT_KEYS], dim=-1))
    if ref.P.use_tp_net and len(e.history_data):
        out["tp_hist"] = torch.stack(list(e.history_data), dim=1)
    return {k: v.detach().clone().float() for k, v in out.items()}


def load_oracle_state(orc: O.HideAndSeekOracle, st):
    for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "cyl", "progress"):
        orc.st[k] = st[k].clone()
    orc.throttle, orc.integ, orc.last_rate = st["throttle"].clone(), st["integ"].clone(), st["last_rate"].clone()
    orc.prev_action, orc.stats = st["prev_action"].clone(), st["stats"].clone()
    orc.tp_hist = st["tp_hist"].clone() if "tp_hist" in st else None


def outputs_from_ref(P, nxt, aux, ref):
    o = dict(state_self=nxt["agents"]["observation"]["state_self"], cylinders=nxt["agents"]["observation"]["cylinders"],
             state_drones=nxt["agents"]["state"]["state_drones"], reward=nxt["agents"]["reward"],
             done=nxt["done"].float(), drone_state=nxt["info"]["drone_state"], prev_action=nxt["info"]["prev_action"],
             cmds=aux["cmds"], ctbr=aux["ctbr"], target_rate=aux["target_rate"], action_error=aux["action_error"],
             stats=torch.cat([nxt["stats"][k] for k in O.STAT_KEYS], dim=-1))
    if P.num_agents > 1:
        o["others"] = nxt["agents"]["observation"]["state_others"]
    if P.use_tp_net:
        tp = nxt["agents"]["TP"]
        o.update(tp_input=tp["TP_input"], tp_groundtruth=tp["TP_groundtruth"], tp_done=tp["TP_done"].float(),
                 tp_pred=ref.env.TP(tp["TP_input"]))
    return {k: v.detach().clone().float() for k, v in o.items()}


def check(name, a, b, rtol=2e-5, atol=2e-6):
    a, b = a.float(), b.float().reshape(a.shape)
    bad = (a - b).abs() > atol + rtol * a.abs()
    if bad.any():
        i = int(torch.argmax((a - b).abs().flatten()))
        raise AssertionError(f"oracle != reference at {name}: {int(bad.sum())}/{bad.numel()} "
                             f"worst ref={a.flatten()[i].item():.7e} oracle={b.flatten()[i].item():.7e}")


def gen_case(name, pk, E, scenario, min_cyl, ticks, progress0):
    torch.manual_seed(0)
    P = O.HSParams(**pk)
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    init = O.sample_reset(P, E, g, scenario, min_cylinders=min_cyl)
    ref = RefEnv(P, E, use_random_cylinder=(scenario == "random_cylinders"),
                 scenario_flag=scenario if scenario != "random_cylinders" else "empty", min_cylinders=min_cyl)
    rec = {}
    if P.use_tp_net:
        for k, v in ref.env.TP.state_dict().items():
            rec[f"tp_weights/{k}"] = v.numpy()
        tp_fn = lambda x: ref.env.TP(x).detach()
    else:
        tp_fn = None
    for k, v in init.items():
        rec[f"init/{k}"] = v.numpy()
    mask = torch.ones(E, dtype=torch.bool)
    r = ref.reset_with(mask, init)
    orc = O.HideAndSeekOracle(P, E)
    o = orc.reset(mask, init, tp_fn)
    ro = dict(state_self=r["agents"]["observation"]["state_self"], cylinders=r["agents"]["observation"]["cylinders"],
              state_drones=r["agents"]["state"]["state_drones"], drone_state=r["info"]["drone_state"])
    if P.use_tp_net:
        ro["tp_input"] = r["agents"]["TP"]["TP_input"]
        ro["tp_pred"] = ref.env.TP(ro["tp_input"]).detach()
    for k, v in ro.items():
        check(f"{name}/reset/{k}", v, o[k])
        rec[f"reset/out/{k}"] = v.detach().clone().float().numpy()
    for k, v in snapshot_state(ref).items():
        rec[f"reset/post/{k}"] = v.numpy()
    if progress0 is not None:
        ref.env.progress_buf[:] = progress0
    done_prev = torch.zeros(E, dtype=torch.bool)
    for t in range(ticks):
        act = torch.randn(E, P.num_agents, 4, generator=g) * (1.5 if t % 2 == 0 else 0.4)
        pre = snapshot_state(ref)
        load_oracle_state(orc, pre)                       # teacher forcing
        if name in UPDATE_EPOCHS:
            ref.env.update_epoch = orc.update_epoch = UPDATE_EPOCHS[name][t]
            rec[f"t{t}/update_epoch"] = np.array(float(UPDATE_EPOCHS[name][t]))
        nxt, aux = ref.step(act, done_prev)
        out = outputs_from_ref(P, nxt, aux, ref)
        want = orc.step(act, done_prev, tp_fn)
        post = snapshot_state(ref)
        for k, v in out.items():
            check(f"{name}/t{t}/{k}", v, want[k].float())
        for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "progress"):
            check(f"{name}/t{t}/post/{k}", post[k], orc.st[k])
        check(f"{name}/t{t}/post/throttle", post["throttle"], orc.throttle)
        check(f"{name}/t{t}/post/integ", post["integ"], orc.integ)
        rec[f"t{t}/action"] = act.numpy()
        rec[f"t{t}/done_prev"] = done_prev.numpy()
        for k, v in pre.items():
            rec[f"t{t}/pre/{k}"] = v.numpy()
        for k, v in out.items():
            rec[f"t{t}/out/{k}"] = v.numpy()
        for k, v in post.items():
            rec[f"t{t}/post/{k}"] = v.numpy()
        done_prev = nxt["done"].reshape(-1).clone()
    rec["meta/params"] = np.array(repr(pk))
    rec["meta/E"], rec["meta/ticks"] = np.array(E), np.array(ticks)
    rec["meta/scenario"] = np.array(scenario)
    rec["meta/min_cyl"] = np.array(min_cyl)
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, f"hs_{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: reference == oracle on {ticks} ticks; wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_partial_reset(name="partial_reset_tp", E=16, warm_ticks=3):
    """IsaacEnv._reset with a PARTIAL `_reset` mask in the middle of an episode, executed by the reference's own source
    (isaac_env.py:210-225 -> hideandseek.py:609-723): the quirks of SURVEY App. D ride on it - the evader's velocity and the
    rate slots of prev_action are NOT reset, first_capture_step is overwritten for every env, the extra physics tick moves
    the envs outside the mask too, and the stats handed back are the pre-reset clone."""
    torch.manual_seed(0)
    P = O.HSParams()
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    init = O.sample_reset(P, E, g, "random_cylinders")
    ref = RefEnv(P, E, use_random_cylinder=True, scenario_flag="empty")
    tp_fn = lambda x: ref.env.TP(x).detach()
    rec = {f"tp_weights/{k}": v.numpy() for k, v in ref.env.TP.state_dict().items()}
    full = torch.ones(E, dtype=torch.bool)
    ref.reset_with(full, init)
    done_prev = torch.zeros(E, dtype=torch.bool)
    for t in range(warm_ticks):
        nxt, _ = ref.step(torch.randn(E, P.num_agents, 4, generator=g), done_prev)
        done_prev = nxt["done"].reshape(-1).clone()
    init2 = O.sample_reset(P, E, g, "random_cylinders")
    mask = torch.rand(E, generator=g) < 0.5
    mask[0], mask[1] = True, False
    pre = snapshot_state(ref)
    orc = O.HideAndSeekOracle(P, E)
    load_oracle_state(orc, pre)
    r = ref.reset_with(mask, init2)
    o = orc.reset(mask, init2, tp_fn)
    ro = dict(state_self=r["agents"]["observation"]["state_self"], cylinders=r["agents"]["observation"]["cylinders"],
              others=r["agents"]["observation"]["state_others"], state_drones=r["agents"]["state"]["state_drones"],
              drone_state=r["info"]["drone_state"], tp_input=r["agents"]["TP"]["TP_input"],
              last_stats=torch.cat([r["stats"][k] for k in O.STAT_KEYS], dim=-1), truncated=r["truncated"].float())
    ro["tp_pred"] = ref.env.TP(ro["tp_input"]).detach()
    post = snapshot_state(ref)
    for k, v in ro.items():
        check(f"{name}/{k}", v, o[k].float())
    for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "progress"):
        check(f"{name}/post/{k}", post[k], orc.st[k])
    check(f"{name}/post/throttle", post["throttle"], orc.throttle)
    check(f"{name}/post/stats", post["stats"], orc.stats)
    check(f"{name}/post/prev_action", post["prev_action"], orc.prev_action)
    rec["mask"] = mask.numpy()
    for k, v in init2.items():
        rec[f"init/{k}"] = v.numpy()
    for k, v in pre.items():
        rec[f"pre/{k}"] = v.numpy()
    for k, v in ro.items():
        rec[f"out/{k}"] = v.detach().clone().float().numpy()
    for k, v in post.items():
        rec[f"post/{k}"] = v.numpy()
    rec["meta/params"] = np.array(repr({}))
    rec["meta/E"] = np.array(E)
    path = os.path.join(OUT_DIR, f"reset_{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: reference == oracle for a partial-mask reset ({int(mask.sum())}/{E} envs); wrote {path}")


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("gen_golden needs the reference tree at /root/reference (build container only)")
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            gen_case(name, *spec)
    if not only or "partial_reset_tp" in only:
        gen_partial_reset()


if __name__ == "__main__":
    main()
