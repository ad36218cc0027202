"""CPU tests of everything that does not need a GPU: the C-ABI library loads and exports every
symbol the header declares, refuses to run without a device (no CPU fallback), the config
composer reproduces the hydra tree, and the tensordict/torchrl stand-ins behave."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    import mupe_b200
    from mupe_b200 import _lib
    header = open(os.path.join(REPO, "include", "hs_b200.h")).read()
    declared = set(re.findall(r"\b(hs_[a-z_]+)\s*\(", header)) - {"hs_default_config"} | {"hs_default_config"}
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/hs_b200.h but not exported"
    assert set(_lib.exported_symbols()) == declared
    assert _lib.lib.hs_abi_version() == _lib.HS_ABI_VERSION


def test_config_struct_layout_matches_c_defaults(built_lib):
    import mupe_b200
    from mupe_b200 import _lib
    c = _lib.default_config(128)            # filled by the C side
    p = mupe_b200.build_hs_config(128)      # filled by the Python side from the YAML/vehicle file
    for name, _ in _lib.hs_config._fields_:
        a, b = getattr(c, name), getattr(p, name)
        if hasattr(a, "__len__"):
            assert list(a) == pytest.approx(list(b), rel=1e-6), name
        else:
            assert a == pytest.approx(b, rel=1e-6), name
    assert c.kf == pytest.approx(0.1259604, rel=1e-6) and c.km == pytest.approx(3.8800789e-3, rel=1e-6)
    assert c.hover_throttle == pytest.approx(0.79548, rel=1e-4) and c.rotor_alpha == pytest.approx(0.4, rel=1e-6)
    n = _lib.lib.hs_arena_floats(ctypes.byref(c))
    assert n == (23 * 3 + 8 + 15) * 128


def test_invalid_configs_are_rejected(built_lib):
    from mupe_b200 import _lib
    for field, bad in (("num_agents", 7), ("num_cylinders", 9), ("obs_max_cylinder", 6), ("num_envs", 0), ("abi_version", 99)):
        c = _lib.default_config(64)
        setattr(c, field, bad)
        assert _lib.lib.hs_arena_floats(ctypes.byref(c)) == -1, field
        assert _lib.lib.hs_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    import mupe_b200
    from mupe_b200 import _lib
    c = _lib.default_config(64)
    h = ctypes.c_void_p()
    assert _lib.lib.hs_create(ctypes.byref(c), ctypes.byref(h)) == -2        # HS_ERR_NO_DEVICE
    assert b"no CPU path" in _lib.lib.hs_last_error()
    with pytest.raises(mupe_b200.HsError):
        mupe_b200.HsEngine(c, "cuda:0")
    cfg = mupe_b200.compose("HideAndSeek", overrides={"task.env.num_envs": 8})
    with pytest.raises(mupe_b200.HsError):
        mupe_b200.HideAndSeek(cfg, headless=True)


def test_compose_reproduces_the_hydra_tree():
    import mupe_b200
    cfg = mupe_b200.compose("HideAndSeek", "mappo", overrides={"task.env.num_envs": 4096, "task.cylinder.max_num": 8})
    assert cfg.task.name == "HideAndSeek" and cfg.env.num_envs == 4096 and cfg.env is cfg.task.env
    assert cfg.env.env_spacing == 5                     # from base/env_base merged @_here_
    assert cfg.sim.dt == 0.01 and cfg.sim.substeps == 1
    assert cfg.task.cylinder.max_num == 8 and cfg.task.cylinder.obs_max_cylinder == 3
    assert cfg.algo.use_TP_net == 1 and cfg.seed == 0
    assert cfg.task.does_not_exist is None              # struct-less OmegaConf behaviour
    g = mupe_b200.compose("HideAndSeek_envgen")
    assert g.task.ratio_unif == 0.3 and g.task.eval_iter == 3
    p = mupe_b200.load_drone_params()
    assert p["mass"] == 0.0321 and p["rotor_configuration"]["time_constant"] == 0.025


def test_tensordict_standin():
    from mupe_b200.compat import TensorDict
    td = TensorDict({"a": torch.zeros(4, 3), "n": {"x": torch.ones(4, 2, 5)}}, [4])
    td[("n", "y")] = torch.arange(4)
    assert td.keys(True, True) == ["a", ("n", "x"), ("n", "y")]
    assert td[("n", "x")].shape == (4, 2, 5) and td.get("zz", None) is None
    sub = td[1:3]
    assert tuple(sub.batch_size) == (2,) and sub[("n", "y")].tolist() == [1, 2]
    td[torch.tensor([0, 2])] = 7.0
    assert td["a"][0, 0] == 7 and td["a"][1, 0] == 0 and td[("n", "x")][2, 0, 0] == 7
    st = torch.stack([td, td], dim=1)
    assert tuple(st.batch_size) == (4, 2) and st["a"].shape == (4, 2, 3) and st.numel() == 8
    c = td.clone()
    c["a"] += 1
    assert td["a"][1, 0] == 0
    sel = td.select("a", ("n", "x"))
    assert sel.keys(True, True) == ["a", ("n", "x")]
    assert td.exclude("a").keys() == ["n"]
    td.update({"n": {"z": torch.zeros(4)}})
    assert ("n", "z") in td and ("n", "x") in td


def test_spec_standins():
    from mupe_b200.compat import BoundedTensorSpec, CompositeSpec, UnboundedContinuousTensorSpec as U
    obs = CompositeSpec({"state_self": U((1, 35)), "state_others": U((2, 3))})
    top = CompositeSpec({"agents": CompositeSpec({"observation": obs.expand(3)})}).expand(16)
    assert tuple(top[("agents", "observation", "state_self")].shape) == (16, 3, 1, 35)
    z = top.zero()
    assert z[("agents", "observation", "state_others")].shape == (16, 3, 2, 3) and tuple(z.batch_size) == (16,)
    act = torch.stack([BoundedTensorSpec(-1, 1, 4)] * 3, dim=0)
    assert tuple(act.shape) == (3, 4)
    top["stats"] = CompositeSpec({"return": U(1)}).expand(16)
    assert ("stats", "return") in top.keys(True, True)


def test_algorithmic_bytes_formula_matches_survey():
    import bench
    ab = bench.algorithmic_bytes(3, 5, 3, 5, 10, True)
    assert ab["total"] == 3077                      # SURVEY.md section 8d / BASELINE.md section 3
    assert ab["tick"] + ab["fill"] == ab["total"]
    assert bench.algorithmic_bytes(3, 8, 3, 5, 10, True)["total"] == 3077 + 36


def test_farthest_point_sampling_and_archive():
    import numpy as np
    from mupe_b200.envs.hideandseek_envgen import GenBuffer, farthest_point_sampling
    pts = torch.tensor([[0.0, 0.0], [0.1, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [0.5, 0.5]])
    idx = farthest_point_sampling(pts, 4).tolist()
    assert idx[0] == 0 and set(idx[1:]) == {2, 3, 4}            # the far corners, never the near-duplicate
    gb = GenBuffer(3, 5, buffer_length=50, rng=np.random.default_rng(0))
    assert gb.task_dim == 27                                      # 18 + 3A in the reference (C = 5)
    rng = np.random.default_rng(1)

    def make_task():
        cells = rng.permutation([(i, j) for i in range(2, 7) for j in range(2, 7)])[:9]
        xy = (cells - 4) * 0.2
        z = np.concatenate([np.full(4, 0.6), np.full(5, 0.6)])
        return np.concatenate([np.concatenate([xy[k], [z[k]]]) for k in range(9)]).astype(np.float32)
    tasks = np.stack([make_task() for _ in range(80)])
    assert all(gb._valid(t) for t in tasks)
    bad = tasks[0].copy(); bad[3:5] = bad[0:2]                    # two drones in one cell
    assert not gb._valid(bad)
    gb.insert_history(tasks[:30])
    assert gb._history_buffer.shape == (30, 27)
    gb.insert_history(tasks[30:])                                 # 80 > 50 -> farthest-point subsample
    assert gb._history_buffer.shape == (50, 27)
    near = gb.samplenearby(16, expand_cylinders=False, expand_step=0.1)
    assert near.shape == (16, 27) and all(gb._valid(t) for t in near)
    assert np.all(np.abs(near[:, 2]) <= 1.3 + 1e-6)
    gb.insert(tasks[:4]); gb.insert_weights(torch.tensor([1.0, 0.0, 1.0, 0.0])); gb.insert_weights(torch.tensor([1.0, 1.0, 0.0, 0.0]))
    gb.update()
    assert gb._weight_buffer.reshape(-1).tolist() == [1.0, 0.5, 0.5, 0.0] and gb._state_buffer.shape == (4, 27)


def test_rollout_and_policy_host_logic(built_lib):
    """Host-side pieces of the rollout rows that need no GPU: stride analysis of compute_gae's inputs,
    parameter-name matching and shapes of the policy wrapper, blob sizing, and the loud failure without a CUDA device."""
    import torch
    import mupe_b200
    from mupe_b200 import _lib, policy, rollout
    # [N, T, k, 1] in the reference's env-major layout and as an [N, T] view of time-major storage
    x = torch.zeros(5, 7, 3, 1)
    assert rollout._col_view(x, "x")[1:] == (3, 21, 3)
    xt = torch.zeros(7, 5, 3, 1).transpose(0, 1)
    assert rollout._col_view(xt, "x")[1:] == (3, 3, 15)
    with pytest.raises(_lib.HsError):
        rollout._col_view(torch.zeros(5, 7, 6)[..., ::2], "x")          # trailing dims must be contiguous
    with pytest.raises(_lib.HsError):
        rollout.compute_gae(x, torch.zeros(5, 7, 1, dtype=torch.bool), x, torch.zeros(5, 3, 1))     # no CPU path
    # parameters under the reference's names, also with the prefixes make_functional / named_parameters give them
    p = policy.init_params(35, 2, 3, 4, True, device="cpu")
    assert p["split_embed.embed.state_self.weight"].shape == (128, 35) and p["attn.in_proj_weight"].shape == (384, 128)
    assert p["head.weight"].shape == (4, 128) and p["log_std"].shape == (4,)
    pref = {("module", "encoder") + tuple(k.split(".")): v for k, v in p.items() if not k.startswith(("head", "log_std"))}
    pref.update({"module.act_dist.fc_mean.weight": p["head.weight"], "module.act_dist.fc_mean.bias": p["head.bias"],
                 "module.act_dist.log_std": p["log_std"]})
    assert policy._find(pref, "attn.out_proj.weight") is p["attn.out_proj.weight"]
    assert policy._find(pref, policy._HEAD["head_w"]) is p["head.weight"]
    assert policy._find(pref, ("log_std",)) is p["log_std"]
    assert policy._find(p, "split_embed.embed.nothing.weight", required=False) is None
    with pytest.raises(_lib.HsError):
        policy.FusedPolicy(p, 2, 3, device="cpu")                        # no CPU path either
    # blob = fp32 part (1 KB aligned) + five tf32 hi/lo weight images; embedding K padded to a multiple of 8
    n35, n20 = _lib.lib.hs_policy_blob_floats(35), _lib.lib.hs_policy_blob_floats(20)
    img = lambda k0: 2 * (k0 // 4) * 2048 + 4 * 2 * 65536
    assert n35 > n20 > 0 and (n35 - n20) * 4 >= img(40) - img(24)
    assert _lib.lib.hs_policy_blob_floats(0) == 0 and _lib.lib.hs_policy_blob_floats(129) == 0


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exactly one JSON line on stdout with
    the contract's keys; it times the reference's own source (or the oracle port) on the host cores and needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("env-steps/sec") and "workload" in d["config"]
    assert d["config"]["ticks_per_step"] == 64 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference-source", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_header_is_plain_c():
    """include/hs_b200.h is the drop-in boundary: it must compile as C99 (no C++ / torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(REPO, "include", "hs_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
