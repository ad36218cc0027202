"""Full-size runs (BASELINE.json configs 2, 3 and the per-launch size of config 4) checked through
size-independent properties, since the CPU oracle cannot replay 65 536 envs in seconds:

* SUB-BATCH EQUIVALENCE: envs are independent, so a random 64-env subset of the big batch, copied into
  a small engine and ticked with the same actions, must reproduce the big batch's rows - bit for bit
  for everything the tick kernel writes, 1e-6 for the predictor-dependent rows (the big batch runs the
  ping-pong tcgen05 kernel, the small one the single-tile kernel; tile neighbours differ, the per-env
  arithmetic does not);
* that same subset against the CPU oracle on a teacher-forced tick (the usual 1e-4 bar);
* invariants: unit quaternions, speed clamp, ground clamp, finite outputs, done == (progress >= max
  length), the chronological TP window shifts by exactly one frame per tick;
* determinism: a second engine fed the same inputs produces identical bytes.
"""
import pytest
import torch

from oracle import hs_oracle as O
from tests.util import assert_close, hs_config_from_params, pull_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FIELDS = ("pos", "quat", "linvel", "angvel", "throttle", "integ", "last_rate", "tpos", "tvel", "progress")


def _field_ids():
    from mupe_b200 import _lib as L
    return dict(pos=L.FIELD_DRONE_POS, quat=L.FIELD_DRONE_ROT, linvel=L.FIELD_DRONE_LINVEL, angvel=L.FIELD_DRONE_ANGVEL,
                throttle=L.FIELD_THROTTLE, integ=L.FIELD_PID_INTEG, last_rate=L.FIELD_PID_LAST_RATE,
                tpos=L.FIELD_TARGET_POS, tvel=L.FIELD_TARGET_VEL, progress=L.FIELD_PROGRESS, cyl=L.FIELD_CYL_POS)


def _device_init(P, E, gen, active_cyl):
    """Random reset poses for E envs, built on the device (the oracle's sampler is too slow at 65 536)."""
    A, C = P.num_agents, P.num_cylinders
    a = P.arena_size / 2 ** 0.5
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    dpos = torch.stack([0.1 + r(E, A) * (a - 0.2), (-a + 0.1) + r(E, A) * (2 * a - 0.2), P.max_height / 2 - 0.1 + 0.2 * r(E, A)], -1)
    tpos = torch.stack([(-a + 0.1) + r(E) * (a - 0.2), (-a + 0.1) + r(E) * (2 * a - 0.2), P.max_height / 2 - 0.1 + 0.2 * r(E)], -1)
    rpy = (r(E, A, 3) - 0.5) * 0.4
    cr, sr, cp, sp, cy, sy = [f(rpy[..., i] * 0.5) for i in range(3) for f in (torch.cos, torch.sin)]
    rot = torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                       cr * cp * sy - sr * sp * cy], -1)
    cyl = torch.zeros(E, C, 3, device=DEV)
    cyl[..., :2] = torch.randint(-3, 4, (E, C, 2), device=DEV, generator=gen).float() * (2 * P.cylinder_size)
    cyl[..., 2] = torch.where(torch.arange(C, device=DEV)[None] < active_cyl, 0.5 * P.max_height, -20.0)
    return dict(drone_pos=dpos, drone_rot=rot, target_pos=tpos, cyl_pos=cyl)


@pytest.mark.parametrize("E,C,active", [(4096, 5, 0), (16384, 8, 8), (65536, 5, 3)],
                         ids=["config2_4096_empty", "config3_16384_8cyl", "config4_65536"])
def test_full_size_properties(E, C, active):
    import mupe_b200
    P = O.HSParams(num_cylinders=C, max_episode_length=40)
    cfg = hs_config_from_params(P, E)
    gen = torch.Generator(device=DEV).manual_seed(E)
    torch.manual_seed(1)
    tp = mupe_b200.TP_net(P.tp_frame_dim, 3 * P.future_step, P.future_step).to(DEV)
    with torch.no_grad():
        for p_ in tp.parameters():
            p_.mul_(2.0)
    init = _device_init(P, E, gen, active)
    big, twin = mupe_b200.HsEngine(cfg, DEV), mupe_b200.HsEngine(cfg, DEV)
    for e in (big, twin):
        e.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        e.step_post_tp(e.tp_weights(tp))
    ids = _field_ids()
    sub = torch.randperm(E, device=DEV, generator=gen)[:64]
    small = mupe_b200.HsEngine(hs_config_from_params(P, 64), DEV)
    small.reset(None, init["drone_pos"][sub], init["drone_rot"][sub], init["target_pos"][sub], init["cyl_pos"][sub])
    small.step_post_tp(small.tp_weights(tp))
    prev_window = None
    T = 42                                          # crosses the episode end at tick 40
    for t in range(T):
        act = torch.randn(E, P.num_agents, 4, device=DEV, generator=gen)
        done_prev = big.out["done"].reshape(E).clone() if t else None
        # ---- sub-batch equivalence: copy the subset's complete state, tick both
        for k in FIELDS + (("cyl",) if C else ()):
            small.set_state(ids[k], big.get_state(ids[k])[sub])
        small.prev_action.copy_(big.prev_action[sub])
        small.stats.copy_(big.stats[:, sub])
        small.v_prey.copy_(big.v_prey)
        small.out["tp_input"].copy_(big.out["tp_input"][sub])
        got = big.step_pre(act, raw=True, reset_pid=done_prev)
        big.step_post_tp(big.tp_weights(tp))
        ref = twin.step_pre(act, raw=True, reset_pid=done_prev)
        twin.step_post_tp(twin.tp_weights(tp))
        s_out = small.step_pre(act[sub].contiguous(), raw=True, reset_pid=None if done_prev is None else done_prev[sub])
        small.step_post_tp(small.tp_weights(tp))
        for k in ("obs_cylinders", "state_others", "reward", "drone_state", "tp_input", "rotor_cmds", "ctbr",
                  "target_rate", "action_error", "done", "tp_groundtruth"):
            assert torch.equal(got[k][sub], s_out[k]), f"t{t}: {k} of a 64-env sub-batch differs from the full batch"
        for k in ("state_self", "state_drones"):
            assert_close(f"t{t}/{k} (sub-batch)", s_out[k], got[k][sub], rtol=1e-6, atol=1e-6)
        # ---- determinism
        for k in ("state_self", "state_drones", "obs_cylinders", "reward", "tp_input", "done"):
            assert torch.equal(got[k], ref[k]), f"t{t}: {k} differs between two engines fed the same inputs"
        # ---- invariants
        st = {k: big.get_state(ids[k]) for k in ("pos", "quat", "linvel", "progress")}
        assert torch.isfinite(got["state_self"]).all() and torch.isfinite(got["reward"]).all()
        assert (st["quat"].norm(dim=-1) - 1).abs().max() < 1e-5
        assert st["linvel"].norm(dim=-1).max() <= P.max_linear_velocity * (1 + 1e-6) if P.max_linear_velocity else True
        assert st["pos"][..., 2].min() >= -1e-6
        assert torch.equal(got["done"].reshape(E), st["progress"] >= P.max_episode_length)
        win = got["tp_input"]
        if prev_window is not None:
            assert torch.equal(win[:, :-1], prev_window[:, 1:]), f"t{t}: TP window did not shift by one frame"
        prev_window = win.clone()
    assert big.out["done"].all()                    # 42 ticks of a 40-tick episode
    # ---- the subset against the CPU oracle on one teacher-forced tick
    orc = O.HideAndSeekOracle(P, 64)
    tp_cpu = mupe_b200.TP_net(P.tp_frame_dim, 3 * P.future_step, P.future_step)
    tp_cpu.load_state_dict({k: v.cpu() for k, v in tp.state_dict().items()})
    tp_fn = lambda x: tp_cpu(x).detach()
    orc.reset(torch.ones(64, dtype=torch.bool), {k: v[sub].cpu() for k, v in init.items()}, tp_fn)
    stt = pull_state(small)
    for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "progress"):
        orc.st[k] = stt[k].clone()
    orc.st["cyl"] = small.get_state(ids["cyl"]).cpu()
    orc.throttle, orc.integ, orc.last_rate = stt["throttle"].clone(), stt["integ"].clone(), stt["last_rate"].clone()
    orc.prev_action = small.prev_action.cpu().clone()
    orc.stats = small.stats.t().cpu().clone()
    orc.tp_hist = small.out["tp_input"].cpu().clone()
    act = torch.randn(64, P.num_agents, 4, device=DEV, generator=gen)
    dprev = small.out["done"].reshape(64).clone()
    from oracle import conditioning as CD
    pre, v_prey = {k: v.clone() for k, v in orc.st.items()}, orc.v_prey
    want = orc.step(act.cpu(), dprev.cpu(), tp_fn)
    cond = CD.TickConditioning(P, v_prey, pre, orc.st, eps=1e-6)
    got = small.step_pre(act, raw=True, reset_pid=dprev)
    small.step_post_tp(small.tp_weights(tp))
    for k, w in (("reward", "reward"), ("drone_state", "drone_state"), ("tp_input", "tp_input"), ("obs_cylinders", "cylinders")):
        cond.check(f"oracle/{k}", got[k], want[w])
    well = (cond.dv < 1e-6) & ~cond.edge          # state_self also carries the (nonlinear) predictor's output
    assert_close("oracle/state_self", got["state_self"].cpu()[well], want["state_self"][well], rtol=2e-4, atol=2e-5)
    for e in (big, twin, small):
        e.close()
