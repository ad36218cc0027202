"""Configuration: a hydra-free composer for the ``cfg/`` tree (same file layout and
``defaults:`` semantics the reference's cfg/train.yaml + cfg/task/*.yaml use) and the
translation of task + vehicle parameters into the C ABI's ``hs_config``.

Reference: cfg/train.yaml:1-40, cfg/task/HideAndSeek.yaml, cfg/base/*.yaml,
omni_drones/robots/drone/multirotor.py:60-76 (yaml.safe_load of the vehicle file).
"""
import math
import os
from typing import Any, Dict, Optional

import numpy as np
import yaml

from . import _lib

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
DEFAULT_CFG_DIR = os.path.join(REPO, "cfg")
CRAZYFLIE_YAML = os.path.join(PKG_DIR, "assets", "crazyflie.yaml")


class Cfg(dict):
    """Attribute-access dict.  Missing keys read as ``None`` -- the reference runs OmegaConf
    with ``set_struct(cfg, False)`` (scripts/train.py:95) and relies on that."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self.get(k, None)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return Cfg({k: Cfg.wrap(v) for k, v in x.items()})
        if isinstance(x, list):
            return [Cfg.wrap(v) for v in x]
        return x

    def pop(self, k, *a):
        return dict.pop(self, k, *a)


def _merge(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _load_group(cfg_dir: str, rel: str) -> dict:
    with open(os.path.join(cfg_dir, rel + ".yaml")) as f:
        body = yaml.safe_load(f) or {}
    out: Dict[str, Any] = {}
    for d in body.pop("defaults", []) or []:
        if isinstance(d, str) and d != "_self_":
            path = d.split("@")[0].lstrip("/")          # "/base/env_base@_here_" -> merge in place
            _merge(out, _load_group(cfg_dir, path))
    return _merge(out, body)


def compose(task: str = "HideAndSeek", algo: str = "mappo", overrides: Optional[dict] = None,
            cfg_dir: str = DEFAULT_CFG_DIR) -> Cfg:
    """Returns the same tree hydra builds for ``train.py task=<task> algo=<algo>``:
    top-level train.yaml keys + ``task`` + ``algo`` + ``sim``/``env`` aliased from the task."""
    root = _load_group(cfg_dir, "train")
    root["task"] = _load_group(cfg_dir, os.path.join("task", task))
    root["algo"] = _load_group(cfg_dir, os.path.join("algo", algo))
    for dotted, v in (overrides or {}).items():
        node = root
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = v
    cfg = Cfg.wrap(root)
    cfg["sim"] = cfg.task.setdefault("sim", Cfg())        # sim: ${task.sim}
    cfg["env"] = cfg.task.setdefault("env", Cfg())        # env: ${task.env}
    return cfg


def load_drone_params(path: str = CRAZYFLIE_YAML) -> dict:
    with open(path) as f:
        return yaml.safe_load(f)


def _f32(x) -> float:
    return float(np.float32(x))


def build_hs_config(num_envs: int, *, num_agents=3, num_cylinders=5, obs_max_cylinder=3,
                    future_step=5, history_step=10, max_episode_length=800, use_tp_net=True,
                    dt=0.01, arena_size=0.9, max_height=1.2, cylinder_size=0.1, catch_radius=0.3,
                    collision_radius=0.07, drone_detect_radius=100.0, target_detect_radius=100.0,
                    v_drone=1.0, mask_value=-5.0, dist_reward_coef=1.0, catch_reward_coef=20.0,
                    detect_reward_coef=0.0, collision_coef=100.0, speed_coef=10.0,
                    smoothness_coef=0.0, smoothness_gated=True, write_smoothness_coef_stat=True,
                    ground_clamp=True, max_linear_velocity=None, use_obstacles=False, contact_mode=0, drone_params: Optional[dict] = None
                    ) -> "_lib.hs_config":
    """All arithmetic on parameters happens here in double precision and is rounded to
    fp32 once, the way the reference's Python scalars meet its fp32 tensors."""
    dp = drone_params or load_drone_params()
    rc = dp["rotor_configuration"]
    rb = dp.get("rigid_body", {})
    c = _lib.default_config(num_envs)
    c.num_agents, c.num_cylinders, c.obs_max_cylinder = num_agents, num_cylinders, obs_max_cylinder
    c.future_step, c.history_step, c.max_episode_length = future_step, history_step, max_episode_length
    c.use_tp_net = int(bool(use_tp_net))
    c.smoothness_gated = int(bool(smoothness_gated))
    c.write_smoothness_coef_stat = int(bool(write_smoothness_coef_stat))
    c.fixed_yaw = int(bool(dp.get("fixed_yaw", 0)))
    c.ground_clamp = int(bool(ground_clamp))
    c.use_obstacles = int(bool(use_obstacles))
    c.contact_mode = int(contact_mode)
    c.drone_radius, c.evader_radius = rb.get("collider_radius", 0.06), 0.05
    c.dt = dt
    c.arena_size, c.max_height, c.cylinder_size = arena_size, max_height, cylinder_size
    c.catch_radius, c.collision_radius = catch_radius, collision_radius
    c.drone_detect_radius, c.target_detect_radius = drone_detect_radius, target_detect_radius
    c.v_drone, c.mask_value = v_drone, mask_value
    c.dist_reward_coef, c.catch_reward_coef, c.detect_reward_coef = dist_reward_coef, catch_reward_coef, detect_reward_coef
    c.collision_coef, c.speed_coef, c.smoothness_coef = collision_coef, speed_coef, smoothness_coef
    c.target_clip, c.max_thrust_ratio = dp["target_clip"], dp["max_thrust_ratio"]
    # rotor constants are formed in fp32 like the reference's parameter tensors (rotor_group.py:42-48)
    w = np.float32(rc["max_rotation_velocities"][0])
    c.kf = float(w * w * np.float32(rc["force_constants"][0]))
    c.km = float(w * w * np.float32(rc["moment_constants"][0]))
    tau = np.clip(np.float32(rc["time_constant"]), np.float32(0), np.float32(1))
    c.rotor_alpha = float(np.float32(dt) / tau)
    for i in range(4):
        c.rotor_dirs[i] = rc["directions"][i]
    xy = rb.get("rotor_link_xy", [[0.028, 0.028], [-0.028, 0.028], [-0.028, -0.028], [0.028, -0.028]])
    for i in range(4):
        c.rotor_x[i], c.rotor_y[i] = xy[i][0], xy[i][1]
    c.drag_coef_times_mass = dp.get("drag_coef", 0.0) * dp["mass"]
    mr = rb.get("rotor_link_mass", 1.0e-4)
    total_mass = dp["mass"] + 4 * mr
    bi = rb.get("base_inertia", [dp["inertia"]["xx"], dp["inertia"]["yy"], dp["inertia"]["zz"]])
    sx = sum(mr * y * y for _, y in xy)
    sy = sum(mr * x * x for x, _ in xy)
    c.total_mass = total_mass
    c.inertia[0], c.inertia[1], c.inertia[2] = bi[0] + sx, bi[1] + sy, bi[2] + sx + sy
    for i in range(3):
        c.inv_inertia[i] = 1.0 / c.inertia[i]
    c.gravity = 9.81
    c.lin_damp_factor = max(0.0, 1.0 - dt * rb.get("linear_damping", 0.2))
    c.ang_damp_factor = max(0.0, 1.0 - dt * rb.get("angular_damping", 0.2))
    vmax = v_drone if max_linear_velocity is None else max_linear_velocity   # hideandseek.py:539
    c.max_linear_velocity = vmax
    c.max_angular_velocity = rb.get("max_angular_velocity", 1000.0)
    c.ground_z = rb.get("collider_half_height", 0.0125)
    g = np.float32(total_mass) * np.float32(9.81)
    c.hover_throttle = float(np.sqrt(g / (np.float32(4) * np.float32(c.kf))))
    c.arena_size_sq = arena_size ** 2
    c.half_arena = 0.5 * arena_size
    c.coll_radius_x2 = 2.0 * collision_radius
    c.vmax_clamped = vmax * (1.0 - 1e-6)
    return c
