"""HideAndSeek_envgen control plane on the device.
CPU: oracle/envgen_oracle.py against the reference's own sanity_check / continuous_to_grid verdicts
(tests/golden/envgen_sanity.npz, oracle/gen_envgen_golden.py) and against the torch FPS restatement.
GPU: hs_fps and hs_gen_sample_nearby bit-exact against that oracle; the env runs a whole
generator cycle with everything resident on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import envgen_oracle as G
from oracle import reset_sampler as RS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "envgen_sanity.npz")


def _history(n, C=5, seed=3, max_height=1.2):
    d = RS.ResetDist.for_task(num_cylinders=C, max_height=max_height, cylinder_height=max_height, seed=seed)
    o = RS.sample_reset(d, n, 1)
    return np.concatenate([o["drone_pos"].reshape(n, -1), o["target_pos"], o["cyl_pos"].reshape(n, -1)], -1)


def test_oracle_acceptance_rule_matches_reference_sanity_check():
    g = np.load(GOLD)
    A, C = int(g["A"]), int(g["C"])
    grid = np.float32(2 * float(g["cylinder_size"]))
    ng = int(float(g["arena"]) * 2 / (2 * float(g["cylinder_size"])))
    inside = G.inside_mask(ng)
    np.testing.assert_array_equal(~inside, g["ref_grid_map"].astype(bool))
    tasks, ref = g["tasks"], g["ref_verdict"]
    assert 0.2 < ref.mean() < 0.8                       # both verdicts well represented
    for t, want, cells in zip(tasks, ref, g["ref_cells"]):
        np.testing.assert_array_equal(G.task_cells(t, A, C, grid, ng), cells)
        assert G.sanity_ok(t, A, C, grid, ng, inside) == bool(want)


def test_oracle_fps_matches_torch_restatement_and_is_greedy():
    import mupe_b200
    from mupe_b200.envs.hideandseek_envgen import farthest_point_sampling
    pts = np.random.default_rng(0).random((700, 27)).astype(np.float32)
    a = G.fps(pts, 60)
    b = farthest_point_sampling(torch.from_numpy(pts), 60).numpy()
    np.testing.assert_array_equal(a, b)                 # generic points: no ties, summation order irrelevant
    # greedy property: each pick maximises the distance to the set chosen so far
    d = ((pts[:, None, :] - pts[None, a, :]) ** 2).sum(-1)
    for i in range(1, 60):
        assert np.isclose(d[a[i], :i].min(), d[:, :i].min(1).max(), rtol=1e-5)


def test_oracle_sample_nearby_properties():
    hist = _history(200)
    r = G.sample_nearby(hist, 300, 3, 5, 0.9, 0.1, 1.2, True, 0.1, seed=5, epoch=1)
    assert 0.3 < r["valid"].mean() <= 1.0
    b = G.task_bounds(3, 5, 0.9, 0.2, 1.2)
    assert (r["tasks"] >= b[:, 0] - 1e-7).all() and (r["tasks"] <= b[:, 1] + 1e-7).all()
    ok = r["valid"].astype(bool)
    delta = r["tasks"][ok][:, :6] - hist[r["origin"][ok]][:, :6]
    assert np.abs(delta[:, [0, 1, 3, 4]]).max() <= 0.1 + 1e-6       # xy noise bounded by expand_step
    cyl = (r["tasks"][ok][:, 12:] - hist[r["origin"][ok]][:, 12:]).reshape(-1, 5, 3)
    assert np.abs(cyl[..., :2]).max() <= 0.2 + 1e-6                  # at most one grid step (clipping may shorten it)
    assert np.abs(cyl[..., 2]).max() == 0.0
    r2 = G.sample_nearby(hist, 300, 3, 5, 0.9, 0.1, 1.2, True, 0.1, seed=5, epoch=2)
    assert not np.array_equal(r["tasks"], r2["tasks"])


# ------------------------------------------------------------------------------------------ GPU
def _gb(A=3, C=5, seed=5, **kw):
    from mupe_b200.envs.hideandseek_envgen import GenBufferDevice
    return GenBufferDevice(A, C, 0.9, 0.1, 1.2, seed=seed, device="cuda:0", **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dim,k,ties", [(3000, 27, 300, False), (3000, 27, 200, True), (257, 3, 257, False),
                                          (70000, 27, 64, False), (400000, 36, 12, False)])
def test_cuda_fps_bit_exact_against_oracle(n, dim, k, ties):
    import mupe_b200  # noqa: F401
    rng = np.random.default_rng(n + k)
    pts = rng.random((n, dim)).astype(np.float32)
    if ties:
        pts = np.round(pts * 2) / 2                     # coarse lattice: many exactly equal distances
    gb = _gb()
    got = gb.fps(torch.from_numpy(pts), k, start=7).cpu().numpy()
    want = G.fps(pts, k, start=7)
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("expand_cyl,step,C", [(True, 0.1, 5), (False, 0.3, 5), (True, 0.5, 8)])
def test_cuda_sample_nearby_bit_exact_against_oracle(expand_cyl, step, C):
    import mupe_b200  # noqa: F401
    hist = _history(400, C=C)
    gb = _gb(C=C, seed=11)
    gb._history_buffer = torch.from_numpy(hist).cuda()
    out, valid = gb.samplenearby(1500, expand_cyl, step, return_valid=True)
    want = G.sample_nearby(hist, 1500, 3, C, 0.9, 0.1, 1.2, expand_cyl, step, seed=11, epoch=gb.epoch)
    np.testing.assert_array_equal(valid.cpu().numpy().astype(np.uint8), want["valid"])
    ok = want["valid"].astype(bool)
    np.testing.assert_array_equal(out.cpu().numpy()[ok], want["tasks"][ok])
    # the public call fills rejected rows with accepted ones: every returned task passes the rule
    full = gb.samplenearby(1500, expand_cyl, step).cpu().numpy()
    inside = G.inside_mask(9)
    assert all(G.sanity_ok(t, 3, C, np.float32(0.2), 9, inside) for t in full[:400])


@pytest.mark.gpu
def test_cuda_sample_nearby_shards_reproduce_the_single_process_stream():
    """task_offset: two ranks drawing 600 tasks each == one process drawing 1200 (same seed, same epoch)."""
    import mupe_b200  # noqa: F401
    hist = _history(300)
    one = _gb(seed=7)
    one._history_buffer = torch.from_numpy(hist).cuda()
    full, valid = one.samplenearby(1200, True, 0.1, return_valid=True)
    parts = []
    for r in range(2):
        gb = _gb(seed=7, task_offset=600 * r)
        gb._history_buffer = torch.from_numpy(hist).cuda()
        parts.append(gb.samplenearby(600, True, 0.1, return_valid=True))
    assert torch.equal(torch.cat([p[1] for p in parts]), valid)
    ok = valid.cpu().numpy()
    np.testing.assert_array_equal(torch.cat([p[0] for p in parts]).cpu().numpy()[ok], full.cpu().numpy()[ok])
    want = G.sample_nearby(hist, 600, 3, 5, 0.9, 0.1, 1.2, True, 0.1, seed=7, epoch=1, task_offset=600)
    np.testing.assert_array_equal(parts[1][1].cpu().numpy().astype(np.uint8), want["valid"])


@pytest.mark.gpu
def test_envgen_runs_a_generator_cycle_on_the_device():
    import mupe_b200 as m
    cfg = m.compose("HideAndSeek_envgen", "mappo", overrides={
        "task.env.num_envs": 512, "task.env.max_episode_length": 6, "task.eval_iter": 2, "task.ratio_unif": 0.5,
        "task.R_min": 0.0, "task.R_max": 1.0, "task.use_random_cylinder": 1, "task.use_particle_generator": 1})
    base = m.IsaacEnv.REGISTRY[cfg.task.name](cfg, headless=True)
    env = m.TransformedEnv(base, m.Compose(m.InitTracker(), m.PIDRateController()))
    try:
        assert base.device_generator
        base.gen_buffer.buffer_length = 300             # force the FPS cap
        td = env.reset()
        for ep in range(5):
            for t in range(6):
                td.set(("agents", "action"), torch.zeros(512, 3, 4, device=base.device))
                td = m.step_mdp(env.step(td))
            td = env.reset()
        hb = base.gen_buffer._history_buffer
        assert hb.is_cuda and 0 < hb.shape[0] <= 300
        assert base.num_unif < 512                      # archive tasks are being replayed
        tasks = base.all_tasks
        assert tasks.is_cuda and tasks.shape == (512, base.gen_buffer.task_dim)
        inside = G.inside_mask(9)
        near = tasks[base.num_unif:].cpu().numpy()
        assert all(G.sanity_ok(t, 3, base.num_cylinders, np.float32(0.2), 9, inside) for t in near[:200])
        assert torch.isfinite(td.get(("agents", "observation", "state_self"))).all()
    finally:
        env.close()
