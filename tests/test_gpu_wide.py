"""GPU: the one-lane-per-env tick (hs_tick_wide_kernel: TMA tensor loads of the SoA state tile, pursuers as a loop).

* For A = 3 it must reproduce the 4-lanes-per-env kernel BIT FOR BIT (same device functions, same operation order):
  outputs, arena, stats - first frame, ragged tiles, PID resets, rotor-command mode, both cylinder capacities, with and
  without the predictor, and through resets.
* For A = 4 (the reference's fixed scenarios carry four pursuer start rows, hideandseek.py:633-682) and A = 5, 6 it is the
  only mapping: parity against the CPU oracle with the conditioning allowances of oracle/conditioning.py, both builds.
"""
import pytest
import torch

from oracle import conditioning as CD
from oracle import hs_oracle as O
from tests.test_gpu_parity import EPS, make_tp, run_case
from tests.util import hs_config_from_params

pytestmark = pytest.mark.gpu
KEYS = ("obs_cylinders", "reward", "done", "drone_state", "rotor_cmds", "ctbr", "target_rate", "action_error", "state_others")


@pytest.mark.parametrize("E,C,tp", [(256, 5, True), (100, 5, True), (32, 8, True), (4, 5, True), (2052, 8, True),
                                    (128, 5, False), (36, 8, False)])
def test_wide_equals_narrow_bit_for_bit(E, C, tp):
    import mupe_b200
    P = O.HSParams(num_cylinders=C, use_tp_net=tp)
    dev = torch.device("cuda:0")
    cfg = hs_config_from_params(P, E)
    g = torch.Generator().manual_seed(E + C)
    init = O.sample_reset(P, E, g, min_cylinders=min(C, 4))
    narrow, wide = mupe_b200.HsEngine(cfg, dev), mupe_b200.HsEngine(cfg, dev)
    narrow.set_tick_mapping(1)
    wide.set_tick_mapping(2)
    keys = KEYS + (("tp_input", "tp_groundtruth", "tp_done") if tp else ("state_self", "state_drones"))

    def same(tag, a, b):
        for k in keys:
            assert torch.equal(a[k], b[k]), f"{tag}: {k} differs"
        # (tile-blocked arena [tile][row][32]; neither kernel touches the padding lanes of a ragged last tile)
        na, wa = narrow.arena.view(-1, 32), wide.arena.view(-1, 32)
        assert torch.equal(na, wa), f"{tag}: arena differs in (tile x row) {torch.nonzero((na != wa).any(1)).flatten().tolist()[:12]}"
        assert torch.equal(narrow.stats, wide.stats), f"{tag}: stats differ"
        assert torch.equal(narrow.prev_action, wide.prev_action), f"{tag}: prev_action differs"

    outs = [e.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"]) for e in (narrow, wide)]
    same("reset", *outs)
    if tp:
        pred = torch.tanh(torch.randn(E, 3 * P.future_step, generator=g)).to(dev)
        for e in (narrow, wide):
            e.step_post(pred)
    for t in range(5):
        act = (torch.randn(E, 3, 4, generator=g) * (1.5 if t % 2 else 0.4)).to(dev)
        rp = (torch.rand(E, generator=g) < 0.3).to(dev) if t == 2 else None
        if t == 3:                                   # rotor commands applied directly (base env without the transform)
            cmds = torch.rand(E, 3, 4, generator=g).to(dev) * 2 - 1
            ae = torch.rand(E, 3, generator=g).to(dev)
            for e in (narrow, wide):
                e.sets[e.next_index()]["action_error"].copy_(ae)
            outs = [e.step_pre(cmds, raw=False) for e in (narrow, wide)]
        else:
            outs = [e.step_pre(act, raw=True, reset_pid=rp) for e in (narrow, wide)]
        same(f"tick {t}", *outs)
        if tp:
            for e in (narrow, wide):
                e.step_post(pred)
    # partial reset in the middle of an episode
    init2 = O.sample_reset(P, E, g, min_cylinders=min(C, 4))
    mask = (torch.rand(E, generator=g) < 0.5).to(dev)
    outs = [e.reset(mask, init2["drone_pos"], init2["drone_rot"], init2["target_pos"], init2["cyl_pos"]) for e in (narrow, wide)]
    same("partial reset", *outs)
    assert torch.equal(outs[0]["truncated"], outs[1]["truncated"])
    outs = [e.step_pre(act, raw=True) for e in (narrow, wide)]
    same("tick after reset", *outs)
    for e in (narrow, wide):
        e.close()


def test_auto_mapping_switches_by_batch_size_and_graphs_replay():
    """auto: 4 lanes per env below 32768 envs, one lane per env from there on; a captured graph of wide ticks replays."""
    import mupe_b200
    P = O.HSParams()
    dev = torch.device("cuda:0")
    E = 32768
    cfg = hs_config_from_params(P, E)
    g = torch.Generator().manual_seed(1)
    init = O.sample_reset(P, E, g)
    torch.manual_seed(0)
    tp = mupe_b200.TP_net(P.tp_frame_dim, 3 * P.future_step, P.future_step).to(dev)
    auto, narrow = mupe_b200.HsEngine(cfg, dev), mupe_b200.HsEngine(cfg, dev)
    narrow.set_tick_mapping(1)
    for e in (auto, narrow):
        e.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        e.step_post_tp(e.tp_weights(tp))
    auto.capture_tick_graphs(auto.tp_weights(tp), raw=True)
    for t in range(3):
        act = torch.randn(E, 3, 4, generator=g).to(dev)
        ref = narrow.step_pre(act, raw=True)
        narrow.step_post_tp(narrow.tp_weights(tp))
        auto.graph_action.copy_(act)
        out = auto.replay_tick()
        for k in ("state_self", "state_drones", "reward", "tp_input", "drone_state", "done"):
            assert torch.equal(out[k], ref[k]), (t, k)
    assert torch.equal(auto.stats, narrow.stats)
    for e in (auto, narrow):
        e.close()


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "ieee"])
def test_four_pursuers_against_oracle(exact):
    run_case(O.HSParams(num_agents=4), E=128, scenario="random_cylinders", steps=10, exact=exact)
    # ('passage' lists the same start pose for pursuers 1 and 3, hideandseek.py:669-676: their mutual downwash is 0/0 in the
    # reference too, so the four-pursuer fixed scenarios used here are the ones with distinct poses)
    run_case(O.HSParams(num_agents=4, num_cylinders=6), E=64, scenario="narrow_gap", steps=8, exact=exact, max_edge_frac=0.3)
    run_case(O.HSParams(num_agents=4, use_tp_net=False), E=36, scenario="wall", steps=6, exact=exact, max_edge_frac=0.3)


@pytest.mark.parametrize("A", [5, 6])
def test_five_and_six_pursuers_against_oracle(A):
    """Beyond the reference's literals (4 start rows): random-cylinder resets only; the oracle is generic in A."""
    run_case(O.HSParams(num_agents=A, num_cylinders=8), E=96, scenario="random_cylinders", steps=8, min_cyl=6)
    run_case(O.HSParams(num_agents=A, use_tp_net=False), E=64, scenario="random_cylinders", steps=6)


def test_wide_three_pursuers_against_oracle_both_builds():
    for exact in (False, True):
        import mupe_b200  # noqa: F401
        P = O.HSParams()
        run_case(P, E=512, scenario="random_cylinders", steps=12, exact=exact, mapping=2)


@pytest.mark.parametrize("mapping", [1, 2], ids=["4-lane", "1-lane"])
def test_analytic_contacts_against_oracle(mapping):
    """hs_config.contact_mode = 1 (cylinder / evader contact response after the integration; the PhysX stand-in's optional
    part, parity unpinned like the integrator): both mappings against oracle.apply_contacts, from states that put pursuers
    INSIDE cylinders and inside the evader's sphere, then through ordinary ticks."""
    import mupe_b200
    from tests.util import push_state, pull_state
    P = O.HSParams(contact_mode=1, num_cylinders=8)
    E, dev = 128, torch.device("cuda:0")
    eng = mupe_b200.HsEngine(hs_config_from_params(P, E), dev)
    eng.set_tick_mapping(mapping)
    orc = O.HideAndSeekOracle(P, E)
    tp_fn = make_tp(P)
    g = torch.Generator().manual_seed(3)
    init = O.sample_reset(P, E, g, min_cylinders=8)
    # pursuer 0 starts 5 cm from the axis of cylinder 0, pursuer 1 3 cm from the evader, pursuer 2 where it was sampled
    init["drone_pos"][:, 0, :2] = init["cyl_pos"][:, 0, :2] + torch.tensor([0.03, 0.04])
    init["drone_pos"][:, 1] = init["target_pos"] + torch.tensor([0.02, -0.02, 0.01])
    mask = torch.ones(E, dtype=torch.bool)
    want = orc.reset(mask, init, tp_fn)
    got = eng.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
    eng.step_post(want["tp_pred"].to(dev))
    st = pull_state(eng)
    d0 = (orc.st["pos"][:, 0, :2] - init["cyl_pos"][:, 0, :2]).norm(dim=-1)
    # projected onto the contact circle of cylinder 0 (a neighbouring cylinder 0.2 m away may push it on: sequential projection)
    assert (torch.isclose(d0, torch.full_like(d0, 0.16), atol=1e-5).float().mean() > 0.6) and (d0 > 0.1).all()
    d1 = (orc.st["pos"][:, 1] - orc.st["tpos"]).norm(dim=-1)
    assert (d1 > 0.1).all()                                                      # pushed out of the evader's sphere (then it moved on)
    for k in ("pos", "linvel"):
        from tests.util import assert_close
        assert_close(f"reset/{k}", st[k], orc.st[k])
    done_prev = torch.zeros(E, dtype=torch.bool)
    for t in range(15):
        act = torch.randn(E, 3, 4, generator=g) * 0.5
        push_state(eng, orc)
        pre, v_prey = {k: v.clone() for k, v in orc.st.items()}, orc.v_prey
        want = orc.step(act, done_prev, tp_fn)
        cond = CD.TickConditioning(P, v_prey, pre, orc.st, eps=EPS[False])
        got = eng.step_pre(act.to(dev), raw=True, reset_pid=done_prev.to(dev))
        eng.step_post(want["tp_pred"].to(dev))
        st = pull_state(eng)
        for k in ("pos", "linvel", "quat", "tpos"):
            cond.check(f"t{t}/state/{k}", st[k], orc.st[k])
        cond.check(f"t{t}/reward", got["reward"], want["reward"])
        cond.check(f"t{t}/drone_state", got["drone_state"], want["drone_state"])
    eng.close()


@pytest.mark.parametrize("E,A,variant", [(256, 3, -1), (100, 3, 0), (128, 3, 1), (300, 3, 2), (2052, 3, 3), (64, 4, None)])
def test_tp_ring_window_equals_shifted_window(E, A, variant):
    """hs_buffers.tp_ring: the frame written twice into a 2H-slot ring gives, as a strided view, the same chronological
    window the plain mode shifts every tick - bit for bit over more than two wraps of the ring, through a partial reset -
    and the fused predictor kernels read it in place (same prediction-dependent rows)."""
    import mupe_b200
    P = O.HSParams(num_agents=A)
    dev = torch.device("cuda:0")
    cfg = hs_config_from_params(P, E)
    g = torch.Generator().manual_seed(E + A)
    init = O.sample_reset(P, E, g)
    plain, ring = mupe_b200.HsEngine(cfg, dev), mupe_b200.HsEngine(cfg, dev)
    plain.set_tick_mapping(2)
    ring.set_tp_ring(True)
    w = None
    if variant is not None:
        torch.manual_seed(0)
        tp = mupe_b200.TP_net(P.tp_frame_dim, 3 * P.future_step, P.future_step).to(dev)
        for e in (plain, ring):
            e.set_predictor_variant(variant)
        w = True
    keys = KEYS + ("tp_groundtruth", "tp_done") + (("state_self", "state_drones") if w else ())

    def post(e):
        if w:
            e.step_post_tp(e.tp_weights(tp))

    def same(tag):
        assert torch.equal(plain.out["tp_input"], ring.tp_window()), f"{tag}: window differs"
        for k in keys:
            assert torch.equal(plain.out[k], ring.out[k]), f"{tag}: {k} differs"
        assert torch.equal(plain.arena, ring.arena) and torch.equal(plain.stats, ring.stats), tag

    for e in (plain, ring):
        e.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        post(e)
    same("reset")
    H = P.history_step
    for t in range(2 * H + 5):
        act = torch.randn(E, A, 4, generator=g).to(dev)
        for e in (plain, ring):
            e.step_pre(act, raw=True)
            post(e)
        same(f"tick {t}")
        if t == H + 2:
            init2 = O.sample_reset(P, E, g)
            mask = (torch.rand(E, generator=g) < 0.5).to(dev)
            for e in (plain, ring):
                e.reset(mask, init2["drone_pos"], init2["drone_rot"], init2["target_pos"], init2["cyl_pos"])
                post(e)
            same("partial reset")
    assert int(ring.tp_ring_pos.min()) == int(ring.tp_ring_pos.max())
    # the ring needs the lane-per-env mapping
    with pytest.raises(mupe_b200.HsError):
        small = mupe_b200.HsEngine(hs_config_from_params(O.HSParams(num_agents=2), 64), dev)
        small.set_tp_ring(True)
    for e in (plain, ring):
        e.close()
