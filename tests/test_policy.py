"""Policy inference next to the tick: the torch-CPU restatement
oracle/policy_oracle.py against tests/golden/policy.npz - produced by the reference's own
PartialAttentionEncoder / DiagGaussian modules (oracle/gen_policy_golden.py) - and, on the GPU, the
fused kernel hs_policy_forward (through mupe_b200.FusedPolicy) against both.  Tolerance: 1e-4
relative fp32 (north_star), absolute floor 2e-5 on O(1) LayerNorm outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import policy_oracle as PO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "policy.npz")
CASES = ["actor_tp", "critic_tp", "actor_notp", "actor_a2_k2"]


def _load(name):
    z = np.load(GOLD)
    p = {k.split("/p/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/p/")}
    obs = {k.split("/obs/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/obs/")}
    ref = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files
           if k.startswith(name + "/ref_") or k == name + "/eps"}
    return p, obs, ref


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_modules(name):
    p, obs, ref = _load(name)
    torch.testing.assert_close(PO.encoder(p, obs), ref["ref_feat"], rtol=1e-5, atol=5e-6)
    if "log_std" in p:
        a, lp, mean = PO.actor(p, obs, ref["eps"])
        torch.testing.assert_close(mean, ref["ref_mean"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(a, ref["ref_action"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lp, ref["ref_logp"], rtol=1e-5, atol=1e-5)
        _, lp0, _ = PO.actor(p, obs, None)
        torch.testing.assert_close(lp0, ref["ref_logp_mode"], rtol=1e-5, atol=1e-5)
    else:
        torch.testing.assert_close(PO.critic(p, obs), ref["ref_value"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 2], ids=["ffma", "tcgen05"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_policy_matches_reference_modules(built_lib, name, impl):
    import mupe_b200
    dev = torch.device("cuda:0")
    p, obs, ref = _load(name)
    pg = {k: v.to(dev).contiguous() for k, v in p.items()}
    og = {k: v.to(dev).contiguous() for k, v in obs.items()}
    net = mupe_b200.FusedPolicy(pg, n_others=obs["state_others"].shape[1], n_cyl=obs["cylinders"].shape[1], device=dev)
    net.impl = impl                     # 1: fp32 FFMA kernel, 2: tcgen05 kernel (3xTF32)
    is_actor = "log_std" in p
    eps = ref["eps"].to(dev).contiguous() if is_actor else None
    out = net(og["state_self"], og["state_others"], og["cylinders"], eps=eps, want_features=True)
    torch.testing.assert_close(out["features"].cpu(), ref["ref_feat"], rtol=1e-4, atol=2e-5)
    if is_actor:
        torch.testing.assert_close(out["head"].cpu(), ref["ref_mean"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["action"].cpu(), ref["ref_action"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["logp"].cpu(), ref["ref_logp"], rtol=1e-4, atol=1e-5)
        mode = net(og["state_self"], og["state_others"], og["cylinders"], eps=None)
        torch.testing.assert_close(mode["action"].cpu(), ref["ref_mean"], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(mode["logp"].cpu(), ref["ref_logp_mode"], rtol=1e-4, atol=1e-5)
    else:
        torch.testing.assert_close(out["head"].cpu(), ref["ref_value"], rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 2], ids=["ffma", "tcgen05"])
@pytest.mark.parametrize("R", [1, 31, 33, 129, 4096 * 3, 148 * 128 + 5])
def test_gpu_policy_ragged_sizes_against_oracle(built_lib, R, impl):
    """Row counts around the 32- and 64-row tiles (both kernels), against the restatement on seeded inputs; refresh()
    picks up in-place parameter updates."""
    import mupe_b200
    dev = torch.device("cuda:0")
    p, _, _ = _load("actor_tp")
    g = torch.Generator().manual_seed(R)
    obs = {"state_self": torch.randn(R, 1, 35, generator=g), "state_others": torch.randn(R, 2, 3, generator=g),
           "cylinders": torch.randn(R, 3, 5, generator=g)}
    eps = torch.randn(R, 4, generator=g)
    pg = {k: v.to(dev).contiguous() for k, v in p.items()}
    net = mupe_b200.FusedPolicy(pg, 2, 3, dev)
    net.impl = impl
    og = {k: v.to(dev) for k, v in obs.items()}
    out = net(og["state_self"], og["state_others"], og["cylinders"], eps=eps.to(dev))
    n = min(R, 3000)                      # the CPU restatement on a prefix and a suffix is enough at the large sizes
    for sl in (slice(0, n), slice(R - n, R)):
        a, lp, mean = PO.actor(p, {k: v[sl] for k, v in obs.items()}, eps[sl])
        torch.testing.assert_close(out["action"][sl].cpu(), a, rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["logp"][sl].cpu(), lp, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        pg["head.bias"].add_(1.0)
    out2 = net.refresh()(og["state_self"], og["state_others"], og["cylinders"], eps=eps.to(dev))
    torch.testing.assert_close(out2["head"], out["head"] + 1.0, rtol=1e-5, atol=1e-5)


def test_oracle_noise_is_standard_normal():
    z = PO.philox_normal(seed=7, step=3, num_rows=50000, head_dim=4)
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs((z[:, 0] * z[:, 1]).mean().item()) < 0.02           # Box-Muller pair uncorrelated
    assert not torch.equal(z, PO.philox_normal(7, 4, 50000, 4))      # the step moves the stream


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 2], ids=["ffma", "tcgen05"])
def test_gpu_policy_in_kernel_noise(built_lib, impl):
    """sample=True: the kernel draws the noise (Philox + Box-Muller) and advances the device step counter itself;
    the noise equals the CPU restatement, the action is mean + std * noise, consecutive calls use consecutive steps,
    and a re-seeded policy replays the same sequence."""
    import mupe_b200
    dev = torch.device("cuda:0")
    p, obs, _ = _load("actor_tp")
    pg = {k: v.to(dev).contiguous() for k, v in p.items()}
    og = {k: v.to(dev).contiguous() for k, v in obs.items()}
    R = obs["state_self"].shape[0]
    net = mupe_b200.FusedPolicy(pg, 2, 3, dev).seed(1234, step=5)
    net.impl = impl
    seq = []
    for k in range(3):
        out = net(og["state_self"], og["state_others"], og["cylinders"], sample=True, want_eps=True)
        want = PO.philox_normal(1234, 5 + k, R, 4)
        torch.testing.assert_close(out["eps"].cpu(), want, rtol=1e-4, atol=1e-5)
        a, lp, _ = PO.actor(p, obs, out["eps"].cpu())
        torch.testing.assert_close(out["action"].cpu(), a, rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["logp"].cpu(), lp, rtol=1e-4, atol=1e-5)
        seq.append(out["action"].clone())
    assert net.rng_state.cpu().tolist() == [1234, 8, 0, 0]
    net.seed(1234, step=5)
    again = net(og["state_self"], og["state_others"], og["cylinders"], sample=True)["action"]
    assert torch.equal(again, seq[0])


@pytest.mark.gpu
@pytest.mark.parametrize("rollout_steps", [0, 6])
def test_gpu_policy_in_the_tick_graph(built_lib, rollout_steps):
    """attach_policy: actor -> critic -> tick -> predictor replayed as ONE CUDA graph per rollout step must equal, bit
    for bit, the same four kernels launched one by one (policy_tick), in ring mode and in rollout-storage mode (where
    action / logp / value land in the time-major rollout rows); and the action it fed to the tick is the actor's."""
    import mupe_b200
    dev = torch.device("cuda:0")
    E, T = 96, 6
    p, _, _ = _load("actor_tp")
    pc, _, _ = _load("critic_tp")
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    g = torch.Generator().manual_seed(2)
    a = 0.9 / 2 ** 0.5
    init = dict(drone_pos=torch.rand(E, 3, 3, generator=g) * 0.4 + torch.tensor([0.1, -0.2, 0.5]),
                drone_rot=torch.tensor([1.0, 0, 0, 0]).expand(E, 3, 4).contiguous(),
                target_pos=torch.rand(E, 3, generator=g) * 0.4 + torch.tensor([-0.5, -0.2, 0.5]),
                cyl_pos=torch.cat([torch.rand(E, 5, 2, generator=g) - 0.5, torch.full((E, 5, 1), 0.6)], -1))
    runs = []
    for use_graph in (False, True):
        eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev, rollout_steps=rollout_steps or None)
        actor = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in p.items()}, 2, 3, dev).seed(99)
        critic = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in pc.items()}, 2, 3, dev)
        w = eng.tp_weights(tp)
        eng.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        eng.step_post_tp(w)
        eng.attach_policy(actor, critic)
        if use_graph:
            eng.capture_tick_graphs(w, raw=True)
        rec = []
        for t in range(T):
            prev = eng.cur
            out = eng.replay_tick() if use_graph else eng.policy_tick(w)
            po = eng.policy_out[eng.cur]
            rec.append({k: out[k].clone() for k in ("state_self", "reward", "drone_state", "tp_input")} |
                       {k: v.clone() for k, v in po.items()})
            if t == 0 and not use_graph:       # the tick consumed the actor's action: same result as feeding it by hand
                obs = eng.sets[prev]
                chk = actor.forward(obs["state_self"], obs["state_others"], obs["obs_cylinders"], eps=None)
                torch.testing.assert_close(chk["head"], po["action_mean"], rtol=0, atol=0)
        if rollout_steps:
            pb = eng.storage.policy_batch()
            assert tuple(pb["action"].shape) == (E, T, 3, 4)
            assert torch.equal(pb["logp"][:, T - 1], rec[-1]["logp"])
        assert eng.launches >= 3 * T              # actor + critic + (tick+predictor as one launch at this batch size)
        runs.append(rec)
        eng.close()
    for t in range(T):
        for k in runs[0][t]:
            assert torch.equal(runs[0][t][k], runs[1][t][k]), f"tick {t}: {k} differs between graph replay and direct launches"
    assert not torch.equal(runs[0][0]["action"], runs[0][1]["action"])
    assert runs[0][0]["state_value"].abs().sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("rollout_steps", [0, 5])
def test_gpu_collector_with_fused_actor_critic(built_lib, rollout_steps):
    """MAPPOActorCritic as the collector's policy (what MAPPOPolicy.__call__ does in the reference's rollout,
    mappo.py:235-251): the [E, T] batch carries action / drone.action_logp / state_value that equal the CPU restatement
    evaluated on the observations of the same batch, with the noise the kernel reports; and the env consumed exactly
    those actions (prev_action = tanh-squashed CTBR of the sampled action, transforms.py:431-441)."""
    import mupe_b200
    dev = torch.device("cuda:0")
    E, T = 48, 5
    cfg = mupe_b200.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": E, "task.sim.device": "cuda:0",
                                                                 "task.env.rollout_steps": rollout_steps})
    base = mupe_b200.IsaacEnv.REGISTRY[cfg.task.name.lower()](cfg, headless=True)
    env = mupe_b200.TransformedEnv(base, mupe_b200.Compose(mupe_b200.InitTracker(), mupe_b200.PIDRateController()))
    p, _, _ = _load("actor_tp")
    pc, _, _ = _load("critic_tp")
    actor = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in p.items()}, 2, 3, dev).seed(5)
    critic = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in pc.items()}, 2, 3, dev)
    pol = mupe_b200.MAPPOActorCritic(actor, critic, keep_noise=True)
    col = mupe_b200.SyncDataCollector(env, policy=pol, frames_per_batch=E * T, total_frames=E * T, return_same_td=True)
    d = next(iter(col)).clone()
    assert tuple(d.batch_size) == (E, T)
    obs = {"state_self": d[("agents", "observation", "state_self")].cpu().flatten(0, 2),
           "state_others": d[("agents", "observation", "state_others")].cpu().flatten(0, 2),
           "cylinders": d[("agents", "observation", "cylinders")].cpu().flatten(0, 2)}
    eps = d["action_noise"].cpu().flatten(0, 2)
    a_w, lp_w, _ = PO.actor(p, obs, eps)
    torch.testing.assert_close(d["drone.action_logp"].cpu().flatten(0, 2), lp_w, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(d["state_value"].cpu().flatten(0, 2), PO.critic(pc, obs), rtol=1e-4, atol=2e-6)
    assert eps.std() > 0.8 and not torch.equal(eps[:E * 3], eps[E * 3:2 * E * 3])       # fresh noise every step
    # ("agents", "action") of the batch holds the rotor commands the PID transform wrote over it (transforms.py:455); the
    # sampled action itself is checked through what the tick derived from it: prev_action = [tanh(a_0..2), clamp((tanh(a_3)+1)/2)]
    prev = d[("next", "info", "prev_action")].cpu().flatten(0, 2)
    want_prev = torch.cat([torch.tanh(a_w[:, :3]), ((torch.tanh(a_w[:, 3:]) + 1) / 2).clamp(0, 0.9)], -1)
    torch.testing.assert_close(prev, want_prev, rtol=1e-4, atol=2e-6)
    # ... and the learner's first step on that batch: GAE straight on the [E, T] views (time-major memory in rollout
    # mode: no copy), bit-identical to the reference's loop on the same numbers (oracle/gae_oracle.py)
    from oracle import gae_oracle as GO
    reward, value, done = d[("next", "agents", "reward")], d["state_value"], d[("next", "done")]
    nobs = d[("next", "agents", "observation")]
    nv = critic.forward(nobs.get("state_self")[:, -1].contiguous(), nobs.get("state_others")[:, -1].contiguous(),
                        nobs.get("cylinders")[:, -1].contiguous())["head"]
    adv, ret = mupe_b200.compute_gae(reward, done, value, nv, 0.995, 0.95)
    a_want, r_want = GO.compute_gae(reward.cpu().numpy(), done.cpu().unsqueeze(2).expand(E, T, 3, 1).numpy(),
                                    value.cpu().numpy(), nv.cpu().numpy(), 0.995, 0.95)
    assert np.array_equal(adv.cpu().numpy(), a_want) and np.array_equal(ret.cpu().numpy(), r_want)
    env.close()


@pytest.mark.gpu
def test_gpu_rollout_graph_with_attached_policy(built_lib):
    """RotatingRolloutGraph with an attached policy: (actor -> critic -> tick) x 4 as ONE graph equals four policy_tick
    calls on a twin engine with the same seed, bit for bit (actions really come from the actor: they change every tick)."""
    import mupe_b200
    from mupe_b200.engine import RotatingRolloutGraph
    dev = torch.device("cuda:0")
    E = 64
    p, _, _ = _load("actor_tp")
    pc, _, _ = _load("critic_tp")
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    g = torch.Generator().manual_seed(3)
    init = dict(drone_pos=torch.rand(E, 3, 3, generator=g) * 0.4 + torch.tensor([0.1, -0.2, 0.5]),
                drone_rot=torch.tensor([1.0, 0, 0, 0]).expand(E, 3, 4).contiguous(),
                target_pos=torch.rand(E, 3, generator=g) * 0.4 + torch.tensor([-0.5, -0.2, 0.5]),
                cyl_pos=torch.cat([torch.rand(E, 5, 2, generator=g) - 0.5, torch.full((E, 5, 1), 0.6)], -1))
    engs = []
    for _ in range(2):
        eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev)
        actor = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in p.items()}, 2, 3, dev).seed(42)
        critic = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in pc.items()}, 2, 3, dev)
        eng.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        eng.step_post_tp(eng.tp_weights(tp))
        eng.attach_policy(actor, critic)
        engs.append(eng)
    rg = RotatingRolloutGraph([engs[0]], [engs[0].tp_weights(tp)], ticks=4)
    acts = []
    for rep in range(2):
        rg.replay()
        for _ in range(4):
            ref = engs[1].policy_tick(engs[1].tp_weights(tp))
            acts.append(engs[1].policy_out[engs[1].cur]["action"].clone())
        for k in ("state_self", "reward", "tp_input", "drone_state"):
            assert torch.equal(engs[0].out[k], ref[k]), (rep, k)
        for k in ("action", "logp", "state_value"):
            assert torch.equal(engs[0].policy_out[engs[0].cur][k], engs[1].policy_out[engs[1].cur][k]), (rep, k)
        assert engs[0].launches == engs[1].launches
    assert not torch.equal(acts[0], acts[1])
    for e in engs:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 2], ids=["ffma", "tcgen05"])
def test_gpu_policy_full_size_properties(built_lib, impl):
    """65 536 envs x 3 agents = 196 608 rows (BASELINE config 4 size), size-independent properties of the network:
    (a) rows are independent - a row permutation of the inputs permutes the outputs; (b) the attention is a set function
    of the keys - swapping the two other-agent tokens, or permuting the cylinder tokens, leaves the result unchanged up to
    fp32 summation order; (c) the two kernels (fp32 FFMA, tcgen05 3xTF32) agree within the parity tolerance."""
    import mupe_b200
    dev = torch.device("cuda:0")
    R = 65536 * 3
    p, _, _ = _load("actor_tp")
    net = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in p.items()}, 2, 3, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    s = torch.randn(R, 1, 35, generator=g, device=dev)
    o = torch.randn(R, 2, 3, generator=g, device=dev)
    c = torch.randn(R, 3, 5, generator=g, device=dev)
    c[torch.rand(R, 3, generator=g, device=dev) < 0.3] = -5.0
    eps = torch.randn(R, 4, generator=g, device=dev)
    base = net(s, o, c, eps=eps, impl=impl, out={})
    perm = torch.randperm(R, generator=g, device=dev)
    out_p = net(s[perm].contiguous(), o[perm].contiguous(), c[perm].contiguous(), eps=eps[perm].contiguous(), impl=impl, out={})
    assert torch.equal(out_p["action"], base["action"][perm]) and torch.equal(out_p["logp"], base["logp"][perm])
    out_s = net(s, o.flip(1).contiguous(), c[:, [2, 0, 1]].contiguous(), eps=eps, impl=impl, out={})
    torch.testing.assert_close(out_s["action"], base["action"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(out_s["logp"], base["logp"], rtol=1e-4, atol=1e-5)
    other = net(s, o, c, eps=eps, impl=3 - impl, out={})
    torch.testing.assert_close(other["action"], base["action"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(other["head"], base["head"], rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
def test_gpu_policy_rollout_graph_with_deferred_critic(built_lib):
    """PolicyRolloutGraph: T x (actor -> tick) plus ONE critic launch over the T stored observations, as one CUDA graph,
    gives the rollout a twin engine produces with actor + critic launched every step: same actions / observations bit
    for bit, the same state values (the critic has no state; batch size may switch its kernel variant: 1e-5), and
    next_state_value = the values GAE bootstraps from."""
    import mupe_b200
    from mupe_b200.engine import PolicyRolloutGraph
    dev = torch.device("cuda:0")
    E, T = 64, 6
    p, _, _ = _load("actor_tp")
    pc, _, _ = _load("critic_tp")
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(16, 15, 5).to(dev)
    g = torch.Generator().manual_seed(5)
    init = dict(drone_pos=torch.rand(E, 3, 3, generator=g) * 0.4 + torch.tensor([0.1, -0.2, 0.5]),
                drone_rot=torch.tensor([1.0, 0, 0, 0]).expand(E, 3, 4).contiguous(),
                target_pos=torch.rand(E, 3, generator=g) * 0.4 + torch.tensor([-0.5, -0.2, 0.5]),
                cyl_pos=torch.cat([torch.rand(E, 5, 2, generator=g) - 0.5, torch.full((E, 5, 1), 0.6)], -1))
    engs = []
    for defer in (True, False):
        eng = mupe_b200.HsEngine(mupe_b200.build_hs_config(E), dev, rollout_steps=T)
        actor = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in p.items()}, 2, 3, dev).seed(7)
        critic = mupe_b200.FusedPolicy({k: v.to(dev).contiguous() for k, v in pc.items()}, 2, 3, dev)
        eng.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        eng.step_post_tp(eng.tp_weights(tp))
        eng.attach_policy(actor, critic, defer_critic=defer)
        engs.append(eng)
    one, ref = engs
    w_ref = ref.tp_weights(tp)
    prg = PolicyRolloutGraph(one, one.tp_weights(tp))            # runs a first rollout eagerly (steady state), then captures
    for _ in range(T):
        ref.policy_tick(w_ref)
    for rep in range(2):
        prg.replay()
        obs0 = {k: ref.sets[ref.cur][k].clone() for k in ("state_self", "state_others", "obs_cylinders")}
        for _ in range(T):
            ref.policy_tick(w_ref)
        torch.cuda.synchronize()
        for k in ("state_self", "reward", "tp_input", "drone_state", "done"):
            assert torch.equal(one.storage.data[k][:T], ref.storage.data[k][:T]), (rep, k)
        for k in ("action", "logp", "action_mean"):
            assert torch.equal(one.storage.policy[k][:T], ref.storage.policy[k][:T]), (rep, k)
        torch.testing.assert_close(one.storage.policy["state_value"][:T], ref.storage.policy["state_value"][:T], rtol=1e-5, atol=1e-6)
        assert one.storage.policy["state_value"][:T].abs().sum() > 0
        # next_state_value[t] = V(observation produced by step t) = state_value[t + 1]; the last one bootstraps GAE
        torch.testing.assert_close(prg.next_state_value[:T - 1], ref.storage.policy["state_value"][1:T], rtol=1e-5, atol=1e-6)
        last = ref.sets[ref.cur]
        want = ref._critic.forward(last["state_self"], last["state_others"], last["obs_cylinders"])["head"]
        torch.testing.assert_close(prg.next_state_value[T - 1], want, rtol=1e-5, atol=1e-6)
        assert one.cur == ref.cur
    for e in engs:
        e.close()
