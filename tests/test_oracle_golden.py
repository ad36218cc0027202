"""CPU: the oracle restatement (oracle/hs_oracle.py) against fixtures produced by the
reference's own source (oracle/gen_golden.py).  Single ticks are replayed from the stored
pre-tick state, so the comparison is tight (2e-5 relative)."""
import pytest
import torch

from oracle import hs_oracle as O
from tests.golden_util import Golden, golden_files, load_oracle_state
from tests.util import assert_close

FILES = golden_files()


def test_fixtures_present():
    assert len(FILES) >= 5, "tests/golden/ is missing fixtures; run `python -m oracle.gen_golden` in the build container"


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("hs_")[-1][:-4])
def test_oracle_replays_reference_ticks(path):
    G = Golden(path)
    P, E = G.P, G.E
    tp_fn = G.tp_fn()
    orc = O.HideAndSeekOracle(P, E)
    init = G.group("init/")
    got = orc.reset(torch.ones(E, dtype=torch.bool), init, tp_fn)
    for k, v in G.group("reset/out/").items():
        assert_close(f"{G.name}/reset/{k}", got[k], v, rtol=2e-5, atol=2e-6)
    for t in range(G.ticks):
        load_oracle_state(orc, G.group(f"t{t}/pre/"))
        act = G.group(f"t{t}/")["action"]
        done_prev = G.group(f"t{t}/")["done_prev"].bool()
        if "update_epoch" in G.group(f"t{t}/"):
            orc.update_epoch = float(G.group(f"t{t}/")["update_epoch"])
        got = orc.step(act, done_prev, tp_fn)
        for k, v in G.group(f"t{t}/out/").items():
            g = got[k].float() if k not in ("done", "tp_done") else got[k].float()
            assert_close(f"{G.name}/t{t}/{k}", g, v, rtol=2e-5, atol=2e-6)
        post = G.group(f"t{t}/post/")
        for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "progress"):
            assert_close(f"{G.name}/t{t}/post/{k}", orc.st[k], post[k], rtol=2e-5, atol=2e-6)
        assert_close(f"{G.name}/t{t}/post/throttle", orc.throttle, post["throttle"], rtol=2e-5, atol=2e-6)
        assert_close(f"{G.name}/t{t}/post/integ", orc.integ, post["integ"], rtol=2e-5, atol=2e-6)


def test_known_answer_constants():
    """Closed-form constants derivable from the reference."""
    P = O.HSParams()
    assert abs(P.kf - 2315.0 ** 2 * 2.350347298350041e-08) < 1e-7
    assert abs(P.km - 2315.0 ** 2 * 7.24e-10) < 1e-9
    orc = O.HideAndSeekOracle(P, 1)
    assert abs(orc.hover_throttle() - 0.79548) < 1e-4
    thrusts, moments, thr = O.rotor_model(P, torch.full((1, 1, 4), 0.2), torch.zeros(1, 1, 4))
    assert abs(float(thr[0, 0, 0]) - 0.4 * (0.6 ** 0.5)) < 1e-6          # one tick from rest, cmd 0.2 -> 0.3098
    assert_close("moment sign", torch.sign(moments[0, 0]), torch.tensor([1., -1., 1., -1.]))


def test_done_tick_divides_stats_once_per_done_step():
    G = Golden([p for p in FILES if p.endswith("hs_done_tick.npz")][0])
    prog = [float(G.group(f"t{t}/post/")["progress"][0]) for t in range(G.ticks)]
    assert prog == [798.0, 799.0, 800.0, 801.0]
    done = [bool(G.group(f"t{t}/out/")["done"][0]) for t in range(G.ticks)]
    assert done == [False, False, True, True]


def test_partial_mask_reset_matches_reference():
    """isaac_env.py:210-225 + hideandseek.py:609-723 run by the reference's own source with a partial `_reset` mask
    (tests/golden/reset_partial_reset_tp.npz, oracle/gen_golden.py::gen_partial_reset)."""
    import os
    import numpy as np
    from tests.golden_util import GOLDEN_DIR, Golden
    G = Golden.__new__(Golden)
    z = np.load(os.path.join(GOLDEN_DIR, "reset_partial_reset_tp.npz"), allow_pickle=False)
    G.z = {k: z[k] for k in z.files}
    G.P, G.E = O.HSParams(), int(G.z["meta/E"])
    tp_fn = G.tp_fn()
    orc = O.HideAndSeekOracle(G.P, G.E)
    pre, post, out = G.group("pre/"), G.group("post/"), G.group("out/")
    load_oracle_state(orc, pre)
    mask = torch.from_numpy(G.z["mask"].copy())
    assert 0 < int(mask.sum()) < G.E
    got = orc.reset(mask, G.group("init/"), tp_fn)
    for k, v in out.items():
        assert_close(f"partial_reset/{k}", got[k].float(), v, rtol=2e-5, atol=2e-6)
    for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "progress"):
        assert_close(f"partial_reset/post/{k}", orc.st[k], post[k], rtol=2e-5, atol=2e-6)
    assert_close("partial_reset/post/stats", orc.stats, post["stats"], rtol=2e-5, atol=2e-6)
    assert_close("partial_reset/post/prev_action", orc.prev_action, post["prev_action"], rtol=2e-5, atol=2e-6)
    # the quirks: evader velocity survives the reset, first_capture_step is rewritten for EVERY env, envs outside the mask
    # still take the extra physics tick
    assert torch.equal(post["tvel"], pre["tvel"])
    assert (post["stats"][:, O.S["first_capture_step"]] == G.P.max_episode_length).all()
    assert not torch.equal(post["pos"][~mask], pre["pos"][~mask])
    assert torch.equal(post["progress"][~mask], pre["progress"][~mask]) and (post["progress"][mask] == 0).all()
