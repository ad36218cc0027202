"""Device-side reset sampler: the CPU restatement oracle/reset_sampler.py
against (a) the published Philox4x32-10 known-answer vectors, (b) tests/golden/reset_grid.npz,
produced by the reference's own rejection_sampling_random_cylinder / grid_to_continuous /
euler_to_quaternion source (oracle/gen_reset_golden.py), and - on the GPU - the CUDA kernel
hs_sample_reset against that restatement (bit-exact for positions, cells and counts)."""
import os

import numpy as np
import pytest

from oracle import reset_sampler as RS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reset_grid.npz")
CASES = {"c5": dict(num_cylinders=5, min_cylinders=0, seed=11),
         "c8": dict(num_cylinders=8, min_cylinders=2, seed=12),
         "c5_big": dict(num_cylinders=5, min_cylinders=0, seed=13, arena_size=1.1)}


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = RS.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert [int(x) for x in got] == want


def _expected_hist(free, C):
    """E[count per cell] when C of the free cells of every env are taken uniformly without replacement."""
    nfree = free.sum(-1, keepdims=True)
    p = free * (C / nfree)
    return p.sum(0), (p * (1 - p)).sum(0)


def check_cell_distribution(cells, free, C, name):
    E = free.shape[0]
    assert np.take_along_axis(free, cells, -1).all(), f"{name}: a cylinder was put on an occupied cell"
    srt = np.sort(cells, -1)
    assert (srt[:, 1:] != srt[:, :-1]).all(), f"{name}: duplicate cells within an env"
    hist = np.bincount(cells.ravel(), minlength=free.shape[1]).astype(np.float64)
    mean, var = _expected_hist(free.astype(np.float64), C)
    z = (hist - mean) / np.sqrt(np.maximum(var, 1e-9))
    z = z[mean > 0]
    assert (hist[mean == 0] == 0).all()
    assert np.abs(z).max() < 5.0, f"{name}: cell histogram off by {np.abs(z).max():.1f} sigma"
    assert abs(z.mean()) < 1.0


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_sampler_matches_reference_grid(case):
    g = np.load(GOLD)
    kw = CASES[case]
    E = int(g[f"{case}/E"])
    d = RS.ResetDist.for_task(**kw)
    o = RS.sample_reset(d, E, epoch=1)
    ng, C = d.num_grid, d.num_cylinders
    # the fixture was generated from these very draws
    np.testing.assert_array_equal(o["drone_pos"], g[f"{case}/drone_pos"])
    np.testing.assert_array_equal(o["target_pos"], g[f"{case}/target_pos"])
    # occupancy grid == the grid_map the reference hands to select_unoccupied_positions (bit-exact)
    np.testing.assert_array_equal(o["occ"].astype(np.int8), g[f"{case}/ref_grid_map"])
    # cell -> metres == grid_to_continuous on every cell
    table = RS.cell_to_xy(d, np.arange(ng * ng))
    np.testing.assert_allclose(table, g[f"{case}/ref_cell_table"], rtol=0, atol=1e-7)
    # quaternion == euler_to_quaternion
    np.testing.assert_allclose(o["drone_rot"], g[f"{case}/ref_rot"], rtol=0, atol=1e-6)
    # distribution: the reference's picks (randperm) and ours against the exact expectation
    free = ~o["occ"].reshape(E, ng * ng)
    ref_xy = g[f"{case}/ref_cyl_xy"]
    ref_cells = np.zeros((E, C), np.int64)
    for k in range(C):
        dist = np.abs(ref_xy[:, k, None, :] - table[None]).sum(-1)
        ref_cells[:, k] = dist.argmin(-1)
        assert dist.min(-1).max() < 1e-6
    check_cell_distribution(ref_cells, free, C, "reference")
    check_cell_distribution(o["cells"], free, C, "oracle")
    # active-cylinder count: uniform on {min..C}, parked cylinders below ground
    lo = kw["min_cylinders"]
    na = o["active_cylinders"].ravel().astype(int)
    assert na.min() == lo and na.max() == C
    for arr in (na, g[f"{case}/ref_active"].ravel().astype(int)):
        h = np.bincount(arr, minlength=C + 1)[lo:]
        exp = E / (C + 1 - lo)
        assert np.abs(h - exp).max() < 5 * np.sqrt(exp)
    z = o["cyl_pos"][..., 2]
    assert ((z == d.cyl_z_inactive) == (np.arange(C)[None] >= na[:, None])).all()
    assert (z[np.arange(C)[None] < na[:, None]] == np.float32(d.cyl_z_active)).all()


def test_draws_depend_only_on_seed_epoch_and_global_index():
    d = RS.ResetDist.for_task(num_cylinders=5, seed=5)
    full = RS.sample_reset(d, 300, epoch=9)
    d2 = RS.ResetDist.for_task(num_cylinders=5, seed=5, env_offset=100)
    part = RS.sample_reset(d2, 120, epoch=9)
    for k in ("drone_pos", "drone_rot", "target_pos", "cyl_pos", "active_cylinders"):
        np.testing.assert_array_equal(part[k], full[k][100:220])
    other = RS.sample_reset(d, 300, epoch=10)
    assert not np.array_equal(other["drone_pos"], full["drone_pos"])
    ranges = full["drone_pos"]
    a = 0.9 / np.sqrt(2.0)
    assert ranges[..., 0].min() >= 0.1 and ranges[..., 0].max() <= a - 0.1 + 1e-6
    assert np.abs(ranges[..., 1]).max() <= a - 0.1 + 1e-6
    assert full["target_pos"][:, 0].max() <= -0.1 + 1e-6


def test_eval_mode_uses_fixed_xy():
    d = RS.ResetDist.for_task(num_cylinders=5, seed=1, use_eval=True)
    o = RS.sample_reset(d, 64, epoch=1)
    assert np.allclose(o["drone_pos"][:, :, :2], np.array([[0.6, 0.0], [0.8, 0.0], [0.8, -0.2]], np.float32))
    assert np.allclose(o["target_pos"][:, :2], [-0.8, 0.0])
    assert np.allclose(o["drone_rot"], [1, 0, 0, 0])


# ------------------------------------------------------------------------------------------ GPU
def _engine(E, C, **cfgkw):
    import torch
    import mupe_b200 as m
    cfg = m.build_hs_config(E, num_agents=3, num_cylinders=C, use_tp_net=False, **cfgkw)
    return m.HsEngine(cfg, torch.device("cuda:0"))


def _dist_struct(d: RS.ResetDist):
    from mupe_b200 import _lib
    s = _lib.hs_reset_dist()
    s.drone_lo[:], s.drone_hi[:] = d.drone_lo, d.drone_hi
    s.target_lo[:], s.target_hi[:] = d.target_lo, d.target_hi
    s.z_lo, s.z_hi = d.z_lo, d.z_hi
    s.rpy_lo[:], s.rpy_hi[:] = d.rpy_lo, d.rpy_hi
    s.grid_size, s.num_grid, s.boundary = d.grid_size, d.num_grid, d.boundary
    s.cyl_z_active, s.cyl_z_inactive = d.cyl_z_active, d.cyl_z_inactive
    s.min_cylinders, s.fixed_num, s.fixed_xy = d.min_cylinders, d.fixed_num, d.fixed_xy
    if d.fixed_xy:
        for k in range(d.num_agents):
            s.fixed_drone_xy[k][0], s.fixed_drone_xy[k][1] = float(d.fixed_drone_xy[k][0]), float(d.fixed_drone_xy[k][1])
        s.fixed_target_xy[:] = [float(x) for x in d.fixed_target_xy]
    s.env_offset, s.seed = d.env_offset, d.seed
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c5", "c8", "c5_big", "fixed_num", "eval", "offset"])
def test_cuda_sampler_bit_exact_against_oracle(case):
    kw = dict(CASES.get(case, dict(num_cylinders=5, seed=21)))
    if case == "fixed_num":
        kw["fixed_num"] = 3
    if case == "eval":
        kw["use_eval"] = True
    if case == "offset":
        kw["env_offset"] = (1 << 32) - 700          # the 32-bit counter word wraps inside the batch
    d = RS.ResetDist.for_task(**kw)
    E = 5000
    eng = _engine(E, d.num_cylinders, arena_size=kw.get("arena_size", 0.9))
    try:
        for epoch in (1, (1 << 40) + 3):
            got = {k: v.cpu().numpy() for k, v in eng.sample_reset(_dist_struct(d), epoch).items()}
            want = RS.sample_reset(d, E, epoch)
            for k in ("drone_pos", "target_pos", "cyl_pos", "active_cylinders"):
                np.testing.assert_array_equal(got[k], want[k], err_msg=f"{case}/{k}")
            np.testing.assert_allclose(got["drone_rot"], want["drone_rot"], rtol=0, atol=1e-6)
    finally:
        eng.close()


@pytest.mark.gpu
def test_cuda_sampler_rejects_impossible_grid():
    import mupe_b200 as m
    d = RS.ResetDist.for_task(num_cylinders=8, arena_size=0.5)      # 5x5 grid: 9 cells inside the circle
    eng = _engine(64, 8)
    try:
        with pytest.raises(m.HsError):
            eng.sample_reset(_dist_struct(d), 1)
    finally:
        eng.close()


@pytest.mark.gpu
def test_env_reset_uses_device_sampler_and_is_reproducible():
    import torch
    import mupe_b200 as m

    def make(seed, **ov):
        cfg = m.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": 256, "task.use_random_cylinder": 1,
                                                           "task.cylinder.max_num": 8, "seed": seed, **ov})
        return m.IsaacEnv.REGISTRY[cfg.task.name](cfg, headless=True)
    e1, e2, e3 = make(3), make(3), make(4)
    try:
        assert e1.device_reset_sampler
        t1, t2, t3 = e1.reset(), e2.reset(), e3.reset()
        k = ("agents", "observation", "cylinders")        # (state_self also carries each env's own TP_net output)
        assert torch.equal(t1.get(k), t2.get(k))
        assert not torch.equal(t1.get(k), t3.get(k))
        f = m._lib.FIELD_DRONE_POS
        assert torch.equal(e1.engine.get_state(f), e2.engine.get_state(f))
        d = RS.ResetDist.for_task(num_cylinders=8, min_cylinders=e1.min_cylinders, seed=3, arena_size=e1.arena_size,
                                  cylinder_size=e1.cylinder_size, max_height=e1.max_height,
                                  cylinder_height=e1.cylinder_height)
        want = RS.sample_reset(d, 256, epoch=1)
        pos = e1.engine.get_state(m._lib.FIELD_CYL_POS).cpu().numpy()
        np.testing.assert_array_equal(pos, want["cyl_pos"])
        np.testing.assert_array_equal(e1.active_cylinders.cpu().numpy(), want["active_cylinders"])
        # second reset draws a new epoch
        t1b = e1.reset()
        assert not torch.equal(t1b.get(k), t1.get(k))
    finally:
        for e in (e1, e2, e3):
            e.close()
