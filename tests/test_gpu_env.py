"""GPU tests of the reference-facing surface: IsaacEnv.REGISTRY -> HideAndSeek(cfg, headless)
-> TransformedEnv(..., PIDRateController) -> reset()/step()/collector, checked against the
CPU oracle driven with the same injected initial state and the same actions.
(Reference call sites: scripts/train.py:110-205, omni_drones/envs/isaac_env.py:210-240.)
"""
import pytest
import torch

from oracle import hs_oracle as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32_lstm():
    # torch lets cuDNN run the LSTM in TF32 by default; the parity bar is fp32
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def make(E, **over):
    import mupe_b200
    o = {"task.env.num_envs": E, "task.sim.device": DEV}
    o.update(over)
    cfg = mupe_b200.compose("HideAndSeek", "mappo", overrides=o)
    base = mupe_b200.IsaacEnv.REGISTRY[cfg.task.name.lower()](cfg, headless=True)
    env = mupe_b200.TransformedEnv(base, mupe_b200.Compose(mupe_b200.InitTracker(), mupe_b200.PIDRateController()))
    return mupe_b200, cfg, base, env


def oracle_for(base, E, **kw):
    P = O.HSParams(num_agents=base.num_agents, num_cylinders=base.num_cylinders,
                   obs_max_cylinder=base.obs_max_cylinder, use_tp_net=base.use_TP_net,
                   max_episode_length=base.max_episode_length, **kw)
    return P, O.HideAndSeekOracle(P, E)


def test_specs_and_registry():
    m, cfg, base, env = make(32)
    assert m.IsaacEnv.REGISTRY["HideAndSeek"] is m.IsaacEnv.REGISTRY["hideandseek"] is m.HideAndSeek
    spec = base.agent_spec["drone"]
    assert spec.n == 3
    assert tuple(spec.observation_spec["state_self"].shape) == (32, 3, 1, 35)
    assert tuple(spec.observation_spec["state_others"].shape) == (32, 3, 2, 3)
    assert tuple(spec.observation_spec["cylinders"].shape) == (32, 3, 3, 5)
    assert tuple(spec.state_spec["state_drones"].shape) == (32, 3, 35)
    assert tuple(spec.action_spec.shape) == (32, 3, 4)
    assert tuple(spec.reward_spec.shape) == (32, 3, 1)
    stats = [k for k in base.observation_spec.keys(True, True) if isinstance(k, tuple) and k[0] == "stats"]
    assert len(stats) == 24
    assert base.drone.params["max_thrust_ratio"] == 0.9 and base.drone.n == 3
    assert isinstance(base.TP, torch.nn.Module)
    with pytest.raises(RuntimeError):
        base.to("cpu")
    env.close()


def test_env_matches_oracle_through_public_api():
    E = 96
    m, cfg, base, env = make(E)
    P, orc = oracle_for(base, E)
    tp_cpu = m.TP_net(P.tp_frame_dim, 3 * P.future_step, P.future_step)
    tp_cpu.load_state_dict({k: v.cpu() for k, v in base.TP.state_dict().items()})
    tp_fn = lambda x: tp_cpu(x).detach()
    torch.manual_seed(5)
    init = base._sample_reset(E)
    cinit = {k: v.cpu() for k, v in init.items()}
    td = env.reset(init=init)
    want = orc.reset(torch.ones(E, dtype=torch.bool), cinit, tp_fn)
    assert td.get("is_init").all() and not td.get("done").any()
    # cuDNN LSTM vs CPU LSTM differ by ~1e-6 in the prediction -> compare at 2e-4
    assert_close("reset/state_self", td[("agents", "observation", "state_self")], want["state_self"], rtol=2e-4, atol=2e-5)
    assert_close("reset/TP_input", td[("agents", "TP", "TP_input")], want["tp_input"])
    g = torch.Generator().manual_seed(11)
    from oracle import conditioning as CD
    traj = CD.TrajectoryConditioning(P, E)           # free-running: strict on the envs that stayed well conditioned
    for t in range(12):
        act = torch.randn(E, 3, 4, generator=g)
        td.set(("agents", "action"), act.to(DEV))
        done_prev = td.get("done").reshape(-1).cpu()
        td = env.step(td)
        pre, v_prey = {k: v.clone() for k, v in orc.st.items()}, orc.v_prey
        want = orc.step(act, done_prev, tp_fn)
        traj.update(v_prey, pre, orc.st)
        nxt = td.get("next")
        # cuDNN-free predictor vs CPU LSTM differ by ~1e-6 in the prediction -> state rows at 2e-4
        traj.check(f"t{t}/state_self", nxt[("agents", "observation", "state_self")], want["state_self"], rtol=2e-4, atol=2e-5)
        traj.check(f"t{t}/state_others", nxt[("agents", "observation", "state_others")], want["others"])
        traj.check(f"t{t}/cylinders", nxt[("agents", "observation", "cylinders")], want["cylinders"])
        assert nxt[("agents", "state", "cylinders")] is nxt[("agents", "observation", "cylinders")]
        traj.check(f"t{t}/state_drones", nxt[("agents", "state", "state_drones")], want["state_drones"], rtol=2e-4, atol=2e-5)
        traj.check(f"t{t}/TP_input", nxt[("agents", "TP", "TP_input")], want["tp_input"])
        traj.check(f"t{t}/reward", nxt[("agents", "reward")], want["reward"])
        traj.check(f"t{t}/drone_state", nxt[("info", "drone_state")], want["drone_state"])
        traj.check(f"t{t}/return", nxt[("stats", "return")], want["stats"][:, O.S["return"]], atol=1e-3)
        # keys the reference's PIDrate transform leaves on the input tensordict
        traj.check(f"t{t}/ctbr", td["ctbr"], want["ctbr"], atol=5e-2)      # rate PID: ~2e4 x rounding of tanh / body rate, free-running
        traj.check(f"t{t}/target_rate", td["target_rate"], want["target_rate"])
        traj.check(f"t{t}/action", td[("agents", "action")], want["cmds"])
        traj.check(f"t{t}/action_error", td[("stats", "action_error_order1")], want["action_error"])
        assert not nxt.get("is_init").any()
        td = m.step_mdp(td)
    assert traj.clean.float().mean() > 0.7, f"only {int(traj.clean.sum())}/{E} envs stayed well conditioned over 12 ticks"
    env.close()


def test_update_epoch_property_reaches_graph_replayed_ticks():
    """scripts/train_deploy.py:270 writes base_env.update_epoch = i; hideandseek.py:988-991 turns it into the smoothness
    coefficient at the next reward call.  The ticks here are CUDA-graph replays captured BEFORE the write."""
    E = 64
    m, cfg, base, env = make(E, **{"task.use_deployment": 1, "task.init_smoothness_coef": 0.5, "task.smooth_lr": 0.4,
                                   "task.max_smoothness_coef": 5.0})
    td = env.reset()
    g = torch.Generator().manual_seed(2)
    coefs, rewards = [], []
    for t, epoch in enumerate([0, 0, 3, 20]):
        base.update_epoch = epoch
        assert base.update_epoch == epoch
        td.set(("agents", "action"), torch.zeros(E, 3, 4, device=DEV))       # same action: only the coefficient changes
        td = env.step(td)
        coefs.append(float(td[("next", "stats", "smoothness_coef")][0]))
        rewards.append(td[("next", "stats", "smoothness_reward")].clone())
        td = m.step_mdp(td)
    assert coefs == pytest.approx([0.5, 0.5, 1.7, 5.0], rel=1e-6), coefs
    assert getattr(base.engine, "_graph_replays", 0) >= 3                   # the fast path really was graph replay
    inc = [rewards[0]] + [rewards[i] - rewards[i - 1] for i in range(1, 4)]
    assert (inc[2] > inc[1] * 2).all() and (inc[3] > inc[2] * 2).all()       # exp(-action_error) x 0.5 -> 1.7 -> 5.0
    env.close()


def test_use_obstacles_puts_the_cylinders_into_the_tp_frame():
    """task.use_obstacles=1 (hideandseek.py:808-817): every TP frame also carries [x, y, size] of the C cylinders, the
    predictor module (input width 7 + 3A + 3C) runs between hs_step_pre and hs_step_post."""
    E = 32
    m, cfg, base, env = make(E, **{"task.use_obstacles": 1})
    C = base.num_cylinders
    assert base.TP.lstm.input_size == 16 + 3 * C
    td = env.reset()
    from mupe_b200 import _lib as L
    cyl = base.engine.get_state(L.FIELD_CYL_POS)
    for t in range(3):
        td.set(("agents", "action"), torch.randn(E, 3, 4, device=DEV))
        td = env.step(td)
        win = td[("next", "agents", "TP", "TP_input")]
        assert tuple(win.shape) == (E, 10, 16 + 3 * C)
        tail = win[:, -1, 16:].reshape(E, C, 3)
        assert torch.equal(tail[..., :2], cyl[..., :2]) and (tail[..., 2] == 0.1).all()
        assert tuple(td[("next", "agents", "observation", "state_self")].shape) == (E, 3, 1, 35)
        td = m.step_mdp(td)
    env.close()


def test_collector_rollout_and_episode_boundary():
    E, T = 64, 8
    m, cfg, base, env = make(E, **{"task.env.max_episode_length": 5})
    frames = []

    def policy(td):
        td.set(("agents", "action"), torch.randn(E, 3, 4, device=DEV))
        return td
    col = m.SyncDataCollector(env, policy=policy, frames_per_batch=E * T, total_frames=E * T * 2, return_same_td=True)
    for data in col:
        frames.append(data)
    assert len(frames) == 2
    d = frames[0]
    assert tuple(d.batch_size) == (E, T)
    assert tuple(d[("next", "agents", "reward")].shape) == (E, T, 3, 1)
    done = d[("next", "done")].reshape(E, T)
    # progress reaches max_episode_length=5 on the 5th tick, every env resets together
    assert done[:, 4].all() and not done[:, :4].any()
    prog = d[("next", "agents", "TP", "TP_input")][:, :, -1, 0]     # newest frame carries raw progress
    assert torch.equal(prog[0].cpu(), torch.tensor([1., 2., 3., 4., 5., 1., 2., 3.]))
    assert d[("is_init")].reshape(E, T)[:, 0].all() and d[("is_init")].reshape(E, T)[:, 5].all()
    # stats returned by reset are the pre-reset values (isaac_env.py:216,223)
    ret_done = d[("next", "stats", "return")].reshape(E, T)[:, 4]
    ret_after_reset = d[("stats", "return")].reshape(E, T)[:, 5]
    assert torch.allclose(ret_done, ret_after_reset)
    assert col._fps > 0
    env.close()


@pytest.mark.parametrize("graph", [0, 1])
def test_rollout_storage_collector_equals_clone_and_stack(graph):
    """with env.rollout_steps=T the tick kernels write `next` straight into the time-major
    RolloutStorage and the collector returns [E, T] views; every entry must equal, bit for bit, what the generic
    clone-per-step + stack collector returns for the same seeds and actions - across a rollout boundary and an
    episode boundary (max_episode_length=5 < T), with and without CUDA graphs."""
    E, T = 64, 8
    outs = []
    for steps in (0, T):
        torch.manual_seed(5)                        # predictor weights
        m, cfg, base, env = make(E, **{"task.env.max_episode_length": 5, "task.env.rollout_steps": steps,
                                       "task.env.cuda_graph": graph})
        g = torch.Generator(device=DEV).manual_seed(3)
        torch.manual_seed(11)                       # reset sampling

        def policy(td):
            td.set(("agents", "action"), torch.randn(E, 3, 4, device=DEV, generator=g))
            return td
        col = m.SyncDataCollector(env, policy=policy, frames_per_batch=E * T, total_frames=E * T * 3, return_same_td=True)
        outs.append([d.clone() for d in col])
        assert (base.engine.storage is not None) == bool(steps)
        env.close()
    assert len(outs[0]) == len(outs[1]) == 3
    for a, b in zip(*outs):
        # (the generic path's step_mdp carries the previous reward along at the root from the second step on;
        # the storage collector fixes its input-side key set at the first step)
        assert set(b.keys(True, True)) <= set(a.keys(True, True))
        assert set(a.keys(True, True)) - set(b.keys(True, True)) <= {("agents", "reward")}
        assert tuple(b.batch_size) == (E, T)
        for k in b.keys(True, True):
            x, y = a.get(k), b.get(k)
            assert x.shape == y.shape, k
            assert torch.equal(x, y), f"{k} differs between the storage collector and clone+stack"


def test_partial_reset_mask_and_direct_rotor_commands():
    E = 32
    import mupe_b200 as m
    cfg = m.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": E, "task.sim.device": DEV,
                                                        "algo.use_TP_net": 0})
    base = m.HideAndSeek(cfg, headless=True)          # no transform: rotor commands go straight in
    td = base.reset()
    assert tuple(td[("agents", "observation", "state_self")].shape) == (E, 3, 1, 20)
    for _ in range(3):
        td.set(("agents", "action"), torch.rand(E, 3, 4, device=DEV) * 2 - 1)
        td = m.step_mdp(base.step(td))
    prog = base.progress_buf
    assert torch.all(prog == 3)
    mask = torch.zeros(E, 1, dtype=torch.bool, device=DEV)
    mask[::2] = True
    td.set("_reset", mask)
    td = base.reset(td)
    prog = base.progress_buf
    assert torch.all(prog[::2] == 0) and torch.all(prog[1::2] == 3)
    base.close()


def test_hover_plumbing_config():
    """BASELINE.json configs[0]: Hover, 64 envs, one Crazyflie -- same fused vehicle tick (A=1, no
    cylinders), checked against the oracle's vehicle stages; obs/reward per hover.py:361-523."""
    import mupe_b200 as m
    E = 64
    cfg = m.compose("Hover", "mappo", overrides={"task.env.num_envs": E, "task.sim.device": DEV})
    base = m.IsaacEnv.REGISTRY["Hover"](cfg, headless=True)
    env = m.TransformedEnv(base, m.Compose(m.InitTracker(), m.PIDRateController()))
    assert tuple(base.agent_spec["drone"].observation_spec.shape) == (E, 1, 20)
    td = env.reset()
    obs0 = td[("agents", "observation")]
    assert obs0.shape == (E, 1, 20)
    P = O.HSParams(num_agents=1, num_cylinders=0, obs_max_cylinder=0, use_tp_net=False, max_episode_length=500,
                   max_linear_velocity=1000.0)
    orc = O.HideAndSeekOracle(P, E)
    from mupe_b200 import _lib as L
    eng = base.engine
    g = torch.Generator().manual_seed(0)
    for t in range(6):
        # teacher-force the oracle's vehicle state from the engine, step both with the same action
        orc.st["pos"], orc.st["quat"] = eng.get_state(L.FIELD_DRONE_POS).cpu(), eng.get_state(L.FIELD_DRONE_ROT).cpu()
        orc.st["linvel"], orc.st["angvel"] = eng.get_state(L.FIELD_DRONE_LINVEL).cpu(), eng.get_state(L.FIELD_DRONE_ANGVEL).cpu()
        orc.throttle, orc.integ = eng.get_state(L.FIELD_THROTTLE).cpu(), eng.get_state(L.FIELD_PID_INTEG).cpu()
        orc.last_rate, orc.prev_action = eng.get_state(L.FIELD_PID_LAST_RATE).cpu(), eng.prev_action.cpu().clone()
        act = torch.randn(E, 1, 4, generator=g)
        c = O.ctbr_pid(P, act, orc.st["quat"], orc.st["angvel"], orc.prev_action, orc.integ, orc.last_rate,
                       torch.zeros(E, dtype=torch.bool))
        thrusts, moments, _ = O.rotor_model(P, c["cmds"], orc.throttle)
        p, q, v, w = O.rigid_body_step(P, orc.st["pos"], orc.st["quat"], orc.st["linvel"], orc.st["angvel"],
                                       thrusts, moments.sum(-1), None)
        td.set(("agents", "action"), act.to(DEV))
        td = env.step(td)
        nxt = td["next"]
        ds = nxt[("info", "drone_state")]
        assert_close(f"t{t}/pos", ds[..., :3], p)
        assert_close(f"t{t}/quat", ds[..., 3:7], q)
        assert_close(f"t{t}/linvel", ds[..., 7:10], v)
        assert_close(f"t{t}/angvel", ds[..., 10:13], w, atol=1e-4)
        obs = nxt[("agents", "observation")]
        assert_close(f"t{t}/obs rpos", obs[..., :3], torch.tensor([0.0, 0.0, 1.0]) - p)
        assert_close(f"t{t}/obs t", obs[..., 16:], torch.full((E, 1, 4), (t + 1) / 500.0))
        pos_err = torch.linalg.vector_norm(torch.tensor([0.0, 0.0, 1.0]) - p, dim=-1)
        up_z = O.quat_basis(q, 2)[..., 2]
        want_r = -10.0 * pos_err + (pos_err <= 0.02).float() * 10 + ((up_z + 1) / 2) ** 2
        far = pos_err > 0.02                                   # heading terms are gated by the position bonus
        assert_close(f"t{t}/reward", nxt[("agents", "reward")][far], want_r[far].unsqueeze(-1), atol=1e-4)
        assert_close(f"t{t}/ctbr", td["ctbr"], c["ctbr"], atol=1e-4 * float(c["ctbr"].abs().max()))
        td = m.step_mdp(td)
    env.close()


@pytest.mark.parametrize("devgen", [1, 0], ids=["device_generator", "host_generator"])
def test_envgen_control_plane(devgen):
    """HideAndSeek_envgen: uniform tasks on a prefix, archive fed every eval_iter episodes
    (hideandseek_envgen.py:875-899, 1302-1333), same tick kernels with the variant flags."""
    import mupe_b200 as m
    E, L = 64, 4
    cfg = m.compose("HideAndSeek_envgen", "mappo", overrides={
        "task.env.num_envs": E, "task.sim.device": DEV, "task.env.max_episode_length": L,
        "task.eval_iter": 2, "task.R_min": 0.0, "task.R_max": 1.0, "task.catch_radius": 5.0,
        "task.env.device_generator": devgen})
    base = m.IsaacEnv.REGISTRY["HideAndSeek_envgen"](cfg, headless=True)
    assert base.device_generator == bool(devgen)
    env = m.TransformedEnv(base, m.Compose(m.InitTracker(), m.PIDRateController()))
    assert ("stats", "ratio_cylinders_5") in base.observation_spec.keys(True, True)
    td = env.reset()
    assert base.num_unif == E and base.gen_buffer._history_buffer.shape[0] == 0
    tasks0 = base.all_tasks.clone() if devgen else base.all_tasks.copy()
    for ep in range(4):
        for t in range(L):
            td.set(("agents", "action"), torch.randn(E, 3, 4, device=DEV) * 0.1)
            td = env.step(td)
            nxt = td["next"]
            td = m.step_mdp(td)
        assert nxt["done"].all()
        # catch_radius 5 m -> every env whose line of sight is not cut by a cylinder succeeds;
        # R_min = 0, R_max = 1 -> every evaluated task is archived after eval_iter = 2 episodes
        assert nxt[("stats", "success")].mean() > 0.5
        if ep == 1:
            assert float(nxt[("stats", "add_history")][0]) == E    # the update tick reports how many tasks were archived ...
        td.set("_reset", nxt["done"].clone())
        td = env.reset(td)
        if ep == 0:
            assert (base.all_tasks == tasks0).all()              # tasks are kept for eval_iter episodes
            assert base.gen_buffer._history_buffer.shape[0] == 0
        if ep == 1:
            assert base.gen_buffer._history_buffer.shape[0] == E   # archive received the E evaluated tasks
            assert base.num_unif == E - int(E * 0.7)               # ratio_unif = 0.3
            # ... and the reset clears it like every other stat (`self.stats[env_ids] = 0.`, hideandseek_envgen.py:997;
            # pinned by tests/test_envgen_episodes.py against the reference's own source)
            assert float(td[("stats", "add_history")][0]) == 0.0
    assert base.gen_buffer._history_buffer.shape[0] >= E
    with pytest.raises(RuntimeError):
        td.set("_reset", torch.zeros(E, 1, dtype=torch.bool, device=DEV))
        env.reset(td)
    env.close()
