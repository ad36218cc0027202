"""Shared helpers for the parity tests: build the C-ABI config from the oracle's parameter
set (so both sides use identical constants) and move state between oracle and engine."""
import numpy as np
import torch

from oracle import hs_oracle as O


def hs_config_from_params(P: O.HSParams, E: int):
    import mupe_b200
    return mupe_b200.build_hs_config(
        E, num_agents=P.num_agents, num_cylinders=P.num_cylinders, obs_max_cylinder=P.obs_max_cylinder,
        future_step=P.future_step, history_step=P.history_step, max_episode_length=P.max_episode_length,
        use_tp_net=P.use_tp_net, dt=P.dt, arena_size=P.arena_size, max_height=P.max_height,
        cylinder_size=P.cylinder_size, catch_radius=P.catch_radius, collision_radius=P.collision_radius,
        drone_detect_radius=P.drone_detect_radius, target_detect_radius=P.target_detect_radius,
        v_drone=P.v_drone, mask_value=P.mask_value, dist_reward_coef=P.dist_reward_coef,
        catch_reward_coef=P.catch_reward_coef, detect_reward_coef=P.detect_reward_coef,
        collision_coef=P.collision_coef, speed_coef=P.speed_coef, smoothness_coef=P.smoothness_coef,
        smoothness_gated=(not P.envgen_variant) and (not P.use_deployment),
        write_smoothness_coef_stat=not P.envgen_variant, ground_clamp=P.ground_clamp,
        max_linear_velocity=P.max_linear_velocity, use_obstacles=P.use_obstacles, contact_mode=P.contact_mode)


def push_state(engine, orc):
    """Copy the oracle's complete state into the engine (teacher forcing)."""
    from mupe_b200 import _lib as L
    st = orc.st
    engine.set_state(L.FIELD_DRONE_POS, st["pos"])
    engine.set_state(L.FIELD_DRONE_ROT, st["quat"])
    engine.set_state(L.FIELD_DRONE_LINVEL, st["linvel"])
    engine.set_state(L.FIELD_DRONE_ANGVEL, st["angvel"])
    engine.set_state(L.FIELD_THROTTLE, orc.throttle)
    engine.set_state(L.FIELD_PID_INTEG, orc.integ)
    engine.set_state(L.FIELD_PID_LAST_RATE, orc.last_rate)
    engine.set_state(L.FIELD_TARGET_POS, st["tpos"])
    engine.set_state(L.FIELD_TARGET_VEL, st["tvel"])
    if engine.C > 0:
        engine.set_state(L.FIELD_CYL_POS, st["cyl"])
    engine.set_state(L.FIELD_PROGRESS, st["progress"])
    engine.prev_action.copy_(orc.prev_action)
    engine.stats.copy_(orc.stats.t().contiguous())
    engine.v_prey.fill_(orc.v_prey)
    if orc.P.use_tp_net and orc.tp_hist is not None:
        engine.out["tp_input"].copy_(orc.tp_hist)


def pull_state(engine):
    from mupe_b200 import _lib as L
    g = lambda f: engine.get_state(f).cpu()
    return dict(pos=g(L.FIELD_DRONE_POS), quat=g(L.FIELD_DRONE_ROT), linvel=g(L.FIELD_DRONE_LINVEL),
                angvel=g(L.FIELD_DRONE_ANGVEL), throttle=g(L.FIELD_THROTTLE), integ=g(L.FIELD_PID_INTEG),
                last_rate=g(L.FIELD_PID_LAST_RATE), tpos=g(L.FIELD_TARGET_POS), tvel=g(L.FIELD_TARGET_VEL),
                progress=g(L.FIELD_PROGRESS))


def assert_close(name, got, want, rtol=1e-4, atol=1e-5):
    """Every element within atol + rtol * |want|; there is deliberately no "fraction of elements may be off" escape
    hatch - ticks that run the evader's sign normalisation or a reward indicator use oracle/conditioning.py, which
    derives per-env allowances from the task's own conditioning."""
    got = got.detach().cpu().float()
    want = want.detach().cpu().float().reshape(got.shape)
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax((err - tol).flatten()))
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.numel()} elements off; "
            f"worst |err|={err.flatten()[i].item():.3e} got={got.flatten()[i].item():.6e} "
            f"want={want.flatten()[i].item():.6e}")
    return 0.0
