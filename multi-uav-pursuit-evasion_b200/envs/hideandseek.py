"""HideAndSeek: 3-pursuer / 1-evader task behind the reference's IsaacEnv surface.

Reference class: omni_drones/envs/hide_and_seek/hideandseek.py:183 (HideAndSeek).
Every per-tick computation of that class (_pre_sim_step, the PhysX step, _compute_state_and_obs,
_compute_reward_and_done) and of the PIDrate action transform runs inside libhs_b200.so; this
module owns what stays on the host side: specs, the tensordict assembly (zero-copy views of the
engine's buffers), reset *sampling* (kept in torch so it can follow the reference's draw order)
and the call of the trajectory predictor between the two halves of a tick.
"""
import math
from typing import Optional

import torch

from .. import _lib
from ..compat import (BoundedTensorSpec, CompositeSpec, TensorDict, Transform, UnboundedContinuousTensorSpec)
from ..config import Cfg, build_hs_config, load_drone_params
from ..engine import HsEngine
from .agent_spec import AgentSpec
from .isaac_env import IsaacEnv
from .tp_net import TP_net

STAT_KEYS = (
    "success", "collision", "blocked", "distance_reward", "distance_predicted_reward", "speed_reward",
    "collision_reward", "collision_wall", "collision_cylinder", "collision_drone", "detect_reward",
    "catch_reward", "smoothness_reward", "smoothness_mean", "smoothness_max", "first_capture_step",
    "sum_detect_step", "return", "action_error_order1_mean", "action_error_order1_max",
    "target_predicted_error", "distance_threshold_L", "out_of_arena", "smoothness_coef",
)

# fixed scenarios: cylinder xy in units of cylinder.size (hideandseek.py:480-531) and start
# poses of up to four pursuers + the evader (hideandseek.py:633-682)
_LAYOUTS = {
    "empty": [],
    "wall": [(0.0, 1.5), (0.0, -1.5), (0.0, 4.5), (0.0, -4.5)],
    "narrow_gap": [(3, -3), (3, 3), (-3, 3), (-3, -3), (0, 3)],
    "random": [(6, 4), (-6, 4), (-2, 4), (0, 2), (-2, -4), (0, -2)],
    "passage": [(0, 3), (-2, 3), (2, 3), (2, -2), (-2, -2), (0, -2)],
}
_STARTS = {
    "empty": ([(0.6, 0.0, 0.5), (0.8, 0.0, 0.5), (0.8, -0.2, 0.5), (0.8, 0.2, 0.5)], (-0.8, 0.0, 0.5)),
    "wall": ([(0.6, 0.4, 0.5), (0.6, 0.0, 0.5), (0.6, -0.4, 0.5), (0.8, 0.2, 0.5)], (-0.8, 0.0, 0.5)),
    "narrow_gap": ([(0.0, 0.7, 0.5), (0.2, 0.7, 0.5), (-0.2, 0.7, 0.5), (0.8, 0.2, 0.5)], (-0.5, 0.2, 0.5)),
    "random": ([(0.6, 0.0, 0.5), (0.8, 0.0, 0.5), (0.8, -0.2, 0.5), (0.8, 0.2, 0.5)], (-0.8, 0.0, 0.5)),
    "passage": ([(0.6, 0.0, 0.5), (0.8, 0.2, 0.5), (0.8, -0.2, 0.5), (0.8, 0.2, 0.5)], (0.0, 0.6, 0.5)),
}


class DroneView:
    """What scripts and transforms read from ``env.drone`` (MultirotorBase, multirotor.py:55-263):
    ``params`` (the vehicle yaml), ``n``, ``action_spec``/``state_spec`` and tensor accessors that
    gather from the engine's arena (they replace the omni.physics.tensors views)."""

    def __init__(self, engine: HsEngine, params: dict, n: int, device):
        self._e, self.params, self.n, self.device = engine, params, n, device
        self.num_rotors = params["rotor_configuration"]["num_rotors"]
        self.action_spec = BoundedTensorSpec(-1, 1, self.num_rotors, device=device)
        self.state_spec = UnboundedContinuousTensorSpec(19 + self.num_rotors, device=device)

    def get_world_poses(self, clone=True):
        return self._e.get_state(_lib.FIELD_DRONE_POS), self._e.get_state(_lib.FIELD_DRONE_ROT)

    def get_velocities(self, clone=True):
        return torch.cat([self._e.get_state(_lib.FIELD_DRONE_LINVEL), self._e.get_state(_lib.FIELD_DRONE_ANGVEL)], -1)

    @property
    def throttle(self):
        return self._e.get_state(_lib.FIELD_THROTTLE)


class HideAndSeek(IsaacEnv):
    VARIANT_ENVGEN = False

    # ------------------------------------------------------------------ construction
    def _design_scene(self):
        t = self.cfg.task
        self.num_agents = int(t.num_agents)
        self.max_cylinders = int(t.cylinder.max_num)
        self.min_cylinders = int(t.cylinder.min_num)
        self.num_cylinders = self.max_cylinders
        self.drone_detect_radius, self.target_detect_radius = t.drone_detect_radius, t.target_detect_radius
        self.catch_radius, self.arena_size, self.max_height = t.catch_radius, t.arena_size, t.max_height
        self.cylinder_size, self.cylinder_height = t.cylinder.size, t.max_height
        self.scenario_flag, self.use_random_cylinder = t.scenario_flag, bool(t.use_random_cylinder)
        self.fixed_num = t.cylinder.fixed_num
        self.use_fixed_num = self.fixed_num is not None
        self.invalid_z, self.boundary = -20.0, t.arena_size - 0.1
        algo = self.cfg.algo
        self.use_TP_net = bool(algo.use_TP_net) if algo is not None and algo.use_TP_net is not None else True
        self.use_eval, self.use_deployment = bool(t.use_eval), bool(t.use_deployment)
        self.obs_max_cylinder = int(t.cylinder.obs_max_cylinder)
        self.future_predcition_step, self.history_step = int(t.future_predcition_step), int(t.history_step)
        self.window_step = int(t.window_step or 1)
        self.use_obstacles = bool(t.use_obstacles)          # TP frame also carries [x, y, size] per cylinder (hideandseek.py:808-817)
        self.time_encoding_dim = 4
        self.collision_radius = t.collision_radius
        self.mask_value = -5
        self._update_epoch = 0
        self.init_smoothness_coef = t.init_smoothness_coef if t.init_smoothness_coef is not None else (t.smoothness_coef or 0.0)
        self.smooth_lr = t.smooth_lr or 0.0
        self.max_smoothness_coef = t.max_smoothness_coef if t.max_smoothness_coef is not None else 5.0
        self.smoothness_coef = min(self.max_smoothness_coef, self.init_smoothness_coef + self.smooth_lr * self._update_epoch)
        if not self.use_random_cylinder and self.scenario_flag not in _LAYOUTS:
            raise ValueError(f"unknown scenario_flag {self.scenario_flag!r}")
        if not self.use_random_cylinder and len(_LAYOUTS[self.scenario_flag]) > self.num_cylinders:
            raise ValueError(f"scenario {self.scenario_flag!r} needs cylinder.max_num >= {len(_LAYOUTS[self.scenario_flag])}")

        params = load_drone_params()
        self._hs_cfg = build_hs_config(
            self.num_envs, num_agents=self.num_agents, num_cylinders=self.num_cylinders,
            obs_max_cylinder=self.obs_max_cylinder, future_step=self.future_predcition_step,
            history_step=self.history_step, max_episode_length=self.max_episode_length,
            use_tp_net=self.use_TP_net, dt=self.dt, arena_size=t.arena_size, max_height=t.max_height,
            cylinder_size=t.cylinder.size, catch_radius=t.catch_radius, collision_radius=t.collision_radius,
            drone_detect_radius=t.drone_detect_radius, target_detect_radius=t.target_detect_radius,
            v_drone=t.v_drone, mask_value=float(self.mask_value), dist_reward_coef=t.dist_reward_coef,
            catch_reward_coef=t.catch_reward_coef, detect_reward_coef=t.detect_reward_coef,
            collision_coef=t.collision_coef, speed_coef=t.speed_coef, smoothness_coef=self.smoothness_coef,
            smoothness_gated=(not self.VARIANT_ENVGEN) and (not self.use_deployment),
            write_smoothness_coef_stat=not self.VARIANT_ENVGEN,
            ground_clamp=True if self.cfg.sim is None or self.cfg.sim.ground_clamp is None else bool(self.cfg.sim.ground_clamp),
            use_obstacles=self.use_obstacles and self.use_TP_net,
            contact_mode=int((self.cfg.sim.contact_mode if self.cfg.sim is not None else 0) or 0),
            max_linear_velocity=t.v_drone, drone_params=params)
        self.engine = HsEngine(self._hs_cfg, self.device, num_output_sets=int(self.cfg.env.output_sets or 2),
                               rollout_steps=int(self.cfg.env.get("rollout_steps", 0) or 0) or None)
        self.v_prey = t.v_drone * t.v_prey                      # hideandseek.py:263
        self.engine.v_prey.fill_(self.v_prey)
        self._curriculum = self.v_prey < 1.3 and not self.VARIANT_ENVGEN
        self.drone = DroneView(self.engine, params, self.num_agents, self.device)
        self.action_is_raw = False      # set by the PIDRateController transform (fused in-kernel)
        self.fused_predictor = True if self.cfg.env.fused_predictor is None else bool(self.cfg.env.fused_predictor)
        self.use_cuda_graph = True if self.cfg.env.cuda_graph is None else bool(self.cfg.env.cuda_graph)
        self._active_fixed = float(len(_LAYOUTS.get(self.scenario_flag, [])))
        # reset poses are drawn by hs_sample_reset (one launch, counter-based stream) unless
        # env.device_reset_sampler=0 selects the torch sampler below (torch's global generator)
        self.device_reset_sampler = True if self.cfg.env.device_reset_sampler is None else bool(self.cfg.env.device_reset_sampler)
        self._reset_epoch = 0
        self._seed = int(self.cfg.seed or 0)
        self._reset_dist = None

        frame = 7 + 3 * self.num_agents + (3 * self.num_cylinders if self.use_obstacles else 0)
        self.TP = TP_net(input_dim=frame, output_dim=3 * self.future_predcition_step,
                         future_predcition_step=self.future_predcition_step, window_step=self.window_step).to(self.device)

    def _set_specs(self):
        n, dev = self.num_agents, self.device
        F, K = self.future_predcition_step, self.obs_max_cylinder
        U = UnboundedContinuousTensorSpec
        D = 3 + (3 * F if self.use_TP_net else 0) + self.time_encoding_dim + 13
        observation_spec = CompositeSpec({"state_self": U((1, D)), "state_others": U((n - 1, 3)), "cylinders": U((K, 5))}).to(dev)
        state_spec = CompositeSpec({"state_drones": U((n, D)), "cylinders": U((K, 5))}).to(dev)
        frame = 7 + 3 * n + (3 * self.num_cylinders if self.use_obstacles else 0)
        TP_spec = CompositeSpec({"TP_input": U((self.history_step, frame)), "TP_groundtruth": U((1, 3)),
                                 "TP_done": U((1, 3))}).to(dev)
        E = self.num_envs
        self.observation_spec = CompositeSpec({"agents": CompositeSpec({
            "observation": observation_spec.expand(n), "state": state_spec, "TP": TP_spec})}).expand(E).to(dev)
        self.action_spec = CompositeSpec({"agents": CompositeSpec({
            "action": torch.stack([self.drone.action_spec] * n, dim=0)})}).expand(E).to(dev)
        self.reward_spec = CompositeSpec({"agents": CompositeSpec({"reward": U((n, 1))})}).expand(E).to(dev)
        self.agent_spec["drone"] = AgentSpec("drone", n, observation_key=("agents", "observation"),
                                             action_key=("agents", "action"), reward_key=("agents", "reward"),
                                             state_key=("agents", "state"))
        stats_spec = CompositeSpec({k: U(1) for k in self._stat_keys()}).expand(E).to(dev)
        info_spec = CompositeSpec({"drone_state": U((n, 13), device=dev),
                                   "prev_action": torch.stack([self.drone.action_spec] * n, 0)}).expand(E).to(dev)
        self.observation_spec["stats"] = stats_spec
        self.observation_spec["info"] = info_spec
        # live views of the engine's buffers (the reference also hands out live references)
        self.stats = TensorDict({k: self.engine.stats[i].unsqueeze(-1) for i, k in enumerate(STAT_KEYS)}, [E], dev)
        self._extra_stats = {}
        self.info = TensorDict({"drone_state": self.engine.out["drone_state"], "prev_action": self.engine.prev_action}, [E], dev)

    def _stat_keys(self):
        return STAT_KEYS

    # ------------------------------------------------------------------ smoothness curriculum
    @property
    def update_epoch(self) -> int:
        return self._update_epoch

    @update_epoch.setter
    def update_epoch(self, i):
        """scripts/train_deploy.py writes ``base_env.update_epoch = i`` every iteration; the reference turns it into
        ``smoothness_coef = min(max, init + smooth_lr * update_epoch)`` at the next reward call (hideandseek.py:988-991).
        The coefficient is a device scalar every tick reads, so already captured CUDA graphs follow it."""
        self._update_epoch = i
        self.smoothness_coef = min(self.max_smoothness_coef, self.init_smoothness_coef + self.smooth_lr * i)
        eng = getattr(self, "engine", None)
        if eng is not None:
            eng.smoothness_coef.fill_(float(self.smoothness_coef))

    # ------------------------------------------------------------------ views
    @property
    def progress_buf(self) -> torch.Tensor:
        return self.engine.get_state(_lib.FIELD_PROGRESS)

    # ------------------------------------------------------------------ reset sampling
    def _uniform(self, lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, device=self.device)

    def _sample_cylinders(self, n, drone_xy, target_xy):
        """Vectorised version of rejection_sampling_random_cylinder + select_unoccupied_positions
        (hideandseek.py:576-607, 106-119): a 9x9 grid of cell size 2*cylinder.size, cells at
        integer distance >= 4 from the centre and the pursuers'/evader's cells are occupied,
        `max_num` distinct free cells drawn without replacement per env (argsort of iid keys
        instead of the reference's per-env CPU randperm loop)."""
        C, cs = self.num_cylinders, self.cylinder_size
        grid_size = 2 * cs
        ng = int(self.arena_size * 2 / grid_size)
        half = int(ng / 2)
        dev = self.device
        ii, jj = torch.meshgrid(torch.arange(ng, device=dev), torch.arange(ng, device=dev), indexing="ij")
        occ = (torch.sqrt(((ii - half) ** 2 + (jj - half) ** 2).float()) >= (ng // 2)).unsqueeze(0).repeat(n, 1, 1)
        cell = lambda xy: torch.clamp(torch.round(xy / grid_size).int() + half, 0, ng - 1).long()
        dc, tc = cell(drone_xy), cell(target_xy)
        ar = torch.arange(n, device=dev)
        for k in range(dc.shape[1]):
            occ[ar, dc[:, k, 0], dc[:, k, 1]] = True
        occ[ar, tc[:, 0, 0], tc[:, 0, 1]] = True
        if self.use_fixed_num:
            n_active = torch.full((n, 1), int(self.fixed_num), device=dev)
        else:
            n_active = torch.randint(self.min_cylinders, C + 1, (n, 1), device=dev)
        keys = torch.rand(n, ng * ng, device=dev)
        keys = torch.where(occ.reshape(n, -1), torch.full_like(keys, 2.0), keys)
        pick = torch.argsort(keys, dim=-1)[:, :C]
        xy = torch.stack([pick // ng, pick % ng], dim=-1).float()
        xy = torch.clamp((xy - half) * grid_size, -self.boundary, self.boundary)
        inactive = torch.arange(C, device=dev).unsqueeze(0) >= n_active
        z = torch.where(inactive, torch.tensor(self.invalid_z, device=dev), torch.tensor(0.5 * self.cylinder_height, device=dev))
        return torch.cat([xy, z.unsqueeze(-1)], dim=-1), n_active.float()

    def _set_seed(self, seed=-1):
        super()._set_seed(seed)
        self._seed = int(seed)
        self._reset_dist = None

    def reset_dist(self) -> "_lib.hs_reset_dist":
        """include/hs_b200.h::hs_reset_dist of this task (hideandseek.py:283-309, 576-598)."""
        if self._reset_dist is None:
            a = self.arena_size / math.sqrt(2.0)
            grid = 2 * self.cylinder_size
            d = _lib.hs_reset_dist()
            d.drone_lo[:], d.drone_hi[:] = [0.1, -a + 0.1], [a - 0.1, a - 0.1]
            d.target_lo[:], d.target_hi[:] = [-a + 0.1, -a + 0.1], [-0.1, a - 0.1]
            d.z_lo, d.z_hi = self.max_height / 2 - 0.1, self.max_height / 2 + 0.1
            if self.use_eval:
                d.rpy_lo[:], d.rpy_hi[:] = [0.0] * 3, [0.0] * 3
                d.fixed_xy = 1
                for k, pnt in enumerate(_STARTS["empty"][0][:self.num_agents]):
                    d.fixed_drone_xy[k][0], d.fixed_drone_xy[k][1] = pnt[0], pnt[1]
                d.fixed_target_xy[:] = list(_STARTS["empty"][1][:2])
            else:
                d.rpy_lo[:], d.rpy_hi[:] = [-0.2 * math.pi, -0.2 * math.pi, 0.0], [0.2 * math.pi] * 3
            d.grid_size, d.num_grid, d.boundary = grid, int(self.arena_size * 2 / grid), self.boundary
            d.cyl_z_active, d.cyl_z_inactive = 0.5 * self.cylinder_height, self.invalid_z
            d.min_cylinders = self.min_cylinders
            d.fixed_num = int(self.fixed_num) if self.use_fixed_num else -1
            d.env_offset = int(self.cfg.env.env_offset or 0)
            d.seed = self._seed & (2 ** 64 - 1)
            self._reset_dist = d
        return self._reset_dist

    def _sample_reset(self, n: int):
        """Initial poses for n envs (draw order follows hideandseek.py:609-697)."""
        A, C, dev = self.num_agents, self.num_cylinders, self.device
        if self.use_random_cylinder and self.device_reset_sampler and n == self.num_envs and C > 0:
            self._reset_epoch += 1
            init = dict(self.engine.sample_reset(self.reset_dist(), self._reset_epoch))
            init["active_cylinders"] = init["active_cylinders"].clone()
            return init
        a = self.arena_size / math.sqrt(2.0)
        zlo, zhi = self.max_height / 2 - 0.1, self.max_height / 2 + 0.1
        if self.use_random_cylinder:
            if not self.use_eval:
                dxy = torch.stack([self._uniform(0.1, a - 0.1, n, A), self._uniform(-a + 0.1, a - 0.1, n, A)], -1)
                txy = torch.stack([self._uniform(-a + 0.1, -0.1, n, 1), self._uniform(-a + 0.1, a - 0.1, n, 1)], -1)
            else:
                dxy = torch.tensor([p[:2] for p in _STARTS["empty"][0][:A]], device=dev).unsqueeze(0).expand(n, -1, -1)
                txy = torch.tensor([_STARTS["empty"][1][:2]], device=dev).unsqueeze(0).expand(n, -1, -1)
            dpos = torch.cat([dxy, self._uniform(zlo, zhi, n, A, 1)], -1)
            tpos = torch.cat([txy, self._uniform(zlo, zhi, n, 1, 1)], -1)
            cyl, n_active = self._sample_cylinders(n, dxy, txy)
        else:
            d, t = _STARTS[self.scenario_flag]
            dpos = torch.tensor(d[:A], device=dev).unsqueeze(0).repeat(n, 1, 1)
            tpos = torch.tensor([t], device=dev).unsqueeze(0).repeat(n, 1, 1)
            cyl = torch.zeros(n, C, 3, device=dev)
            cyl[..., 0] = torch.arange(C, device=dev) * 2 * self.cylinder_size
            cyl[..., 2] = self.invalid_z
            for k, (x, y) in enumerate(_LAYOUTS[self.scenario_flag]):
                cyl[:, k] = torch.tensor([x * self.cylinder_size, y * self.cylinder_size, 0.5 * self.cylinder_height], device=dev)
            n_active = torch.full((n, 1), self._active_fixed, device=dev)
        if self.use_eval:
            rpy = torch.zeros(n, A, 3, device=dev)
        else:
            lo = torch.tensor([-0.2, -0.2, 0.0], device=dev) * torch.pi
            hi = torch.tensor([0.2, 0.2, 0.2], device=dev) * torch.pi
            rpy = lo + (hi - lo) * torch.rand(n, A, 3, device=dev)
        r, p, y = rpy.unbind(-1)
        cy, sy, cp, sp, cr, sr = torch.cos(y * 0.5), torch.sin(y * 0.5), torch.cos(p * 0.5), torch.sin(p * 0.5), \
            torch.cos(r * 0.5), torch.sin(r * 0.5)
        rot = torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                           cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], dim=-1)
        return dict(drone_pos=dpos, drone_rot=rot, target_pos=tpos.squeeze(1), cyl_pos=cyl, active_cylinders=n_active)

    # ------------------------------------------------------------------ tensordict assembly
    def _obs_td(self, out) -> TensorDict:
        E, dev = self.num_envs, self.device
        obs = TensorDict({"state_self": out["state_self"], "cylinders": out["obs_cylinders"]}, [E, self.num_agents], dev)
        if self.num_agents > 1:
            obs.set("state_others", out["state_others"])
        state = TensorDict({"state_drones": out["state_drones"], "cylinders": out["obs_cylinders"]}, [E], dev)
        agents = {"observation": obs, "state": state}
        if self.use_TP_net:
            agents["TP"] = TensorDict({"TP_input": out["tp_input"], "TP_groundtruth": out["tp_groundtruth"],
                                       "TP_done": out["tp_done"]}, [E], dev)
        self.info.set("drone_state", out["drone_state"])
        return TensorDict({"agents": agents, "stats": self.stats, "info": self.info}, [E], dev)

    def rollout_next_td(self, stats: torch.Tensor, prev_action: torch.Tensor) -> TensorDict:
        """``next`` of a finished rollout as ``[E, T]`` views of the engine's time-major RolloutStorage
        (rollout mode, ``env.rollout_steps=T``): nothing is copied.  ``stats`` ``[T, 24, E]`` and
        ``prev_action`` ``[T, E, A, 4]`` are the per-step snapshots of the two live buffers the tick
        updates in place (the reference's collector clones them with every step too)."""
        st = self.engine.storage
        if st is None:
            raise RuntimeError("rollout_next_td needs rollout mode (cfg.env.rollout_steps)")
        E, T, A, dev = self.num_envs, st.T, self.num_agents, self.device
        b = st.batch()
        obs = TensorDict({"state_self": b["state_self"], "cylinders": b["obs_cylinders"]}, [E, T, A], dev)
        if A > 1:
            obs.set("state_others", b["state_others"])
        state = TensorDict({"state_drones": b["state_drones"], "cylinders": b["obs_cylinders"]}, [E, T], dev)
        agents = {"observation": obs, "state": state, "reward": b["reward"]}
        if self.use_TP_net:
            agents["TP"] = TensorDict({"TP_input": b["tp_input"], "TP_groundtruth": b["tp_groundtruth"],
                                       "TP_done": b["tp_done"]}, [E, T], dev)
        stats_td = TensorDict({k: stats[:, i].transpose(0, 1).unsqueeze(-1) for i, k in enumerate(STAT_KEYS)}, [E, T], dev)
        info = TensorDict({"drone_state": b["drone_state"], "prev_action": prev_action.transpose(0, 1)}, [E, T], dev)
        return TensorDict({"agents": agents, "stats": stats_td, "info": info, "done": b["done"]}, [E, T], dev)

    def _predict(self, out):
        """Second half of the tick.  A TP_net-shaped predictor is evaluated inside the fused
        kernel from its live parameters; any other module goes through torch and hs_step_post
        (same arithmetic downstream, the module's own forward upstream)."""
        if not self.use_TP_net:
            return
        w = self.engine.tp_weights(self.TP) if self.fused_predictor else None
        if w is not None:
            self.engine.step_post_tp(w)
        else:
            with torch.no_grad():
                pred = self.TP(out["tp_input"])
            self.engine.step_post(pred)

    # ------------------------------------------------------------------ EnvBase protocol
    def _reset(self, tensordict: Optional[TensorDict] = None, init: Optional[dict] = None, **kwargs) -> TensorDict:
        E, dev = self.num_envs, self.device
        if tensordict is not None and "_reset" in tensordict:
            mask = tensordict.get("_reset").reshape(E)
        else:
            mask = None
        last_stats = self.stats.clone()
        if init is None:
            init = self._sample_reset(E)          # rows outside the mask are ignored by the kernel
        if "active_cylinders" in init:
            ac = init["active_cylinders"]
            self.active_cylinders = ac if mask is None or not hasattr(self, "active_cylinders") \
                else torch.where(mask.unsqueeze(-1), ac, self.active_cylinders)
        out = self.engine.reset(mask, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
        self._predict(out)
        td = TensorDict({}, self.batch_size, dev)
        td.update(self._obs_td(out))
        td.set("stats", last_stats)
        td.set("truncated", out["truncated"])
        return td

    def _graph_ready(self) -> bool:
        """(Re)captures the per-tick CUDA graphs when the fast path applies: raw actions through
        the fused PID, and either no predictor or a TP_net-shaped one whose parameters still
        live where they did at capture time."""
        if not (self.use_cuda_graph and self.action_is_raw):
            return False
        eng = self.engine
        w = None
        if self.use_TP_net:
            w = eng.tp_weights(self.TP) if self.fused_predictor else None
            if w is None:
                return False
        key = None if w is None else (w.weight_ih, w.weight_hh, w.bias_ih, w.bias_hh, w.fc_weight, w.fc_bias)
        if getattr(self, "_graph_key", ...) != key or getattr(eng, "_graphs", None) is None:
            eng.capture_tick_graphs(w, raw=True)
            self._graph_key = key
        return True

    def _step(self, tensordict: TensorDict) -> TensorDict:
        action = tensordict.get(("agents", "action"))
        eng = self.engine
        if self.action_is_raw:
            done_prev = tensordict.get("done", None)
            if self._graph_ready():
                eng.graph_action.copy_(action.reshape(eng.graph_action.shape))
                if done_prev is None:
                    eng.graph_reset_pid.zero_()
                else:
                    eng.graph_reset_pid.copy_(done_prev.reshape(-1))
                out = eng.replay_tick()
            else:
                out = eng.step_pre(action.contiguous(), raw=True, reset_pid=done_prev)
                self._predict(out)
            # keys the reference's transform writes on the input tensordict (transforms.py:441-458)
            tensordict.set(("stats", "action_error_order1"), out["action_error"])
            tensordict.set(("info", "prev_action"), eng.prev_action)
            tensordict.set(("agents", "action"), out["rotor_cmds"])
            tensordict.set("ctbr", out["ctbr"])
            tensordict.set("target_rate", out["target_rate"])
        else:
            ae = tensordict.get(("stats", "action_error_order1"), None)
            if ae is None:
                ae = torch.zeros(self.num_envs, self.num_agents, device=self.device)
            eng.sets[eng.next_index()]["action_error"].copy_(ae)
            out = eng.step_pre(action.contiguous(), raw=False, reset_pid=None)
            self._predict(out)
        nxt = self._obs_td(out)
        nxt.set(("agents", "reward"), out["reward"])
        nxt.set("done", out["done"])
        # evader-speed curriculum (hideandseek.py:1012-1015): its gate `torch.any(done)` can only open on a tick where some
        # env finishes an episode - the host-side progress bound knows those ticks, so the two all_reduces of a sharded job
        # run once per episode instead of every tick; the flag is dropped once the speed has reached its cap
        if self._curriculum and eng.maybe_done():
            self._update_v_prey(out["done"])
            if float(eng.v_prey) >= 1.3:
                self._curriculum = False
        return TensorDict({"next": nxt}, self.batch_size, self.device)

    def _update_v_prey(self, done):
        """hideandseek.py:1012-1015 without a host sync: v_prey lives in device memory.  In a
        sharded job the success rate and the done flag are reduced over all ranks."""
        from ..parallel import curriculum_step
        curriculum_step(self.engine.v_prey, done, self.engine.stats[0])

    def gather_episode_returns(self) -> torch.Tensor:
        """Per-env episode returns of the whole job in global env order (one all_gather)."""
        from ..parallel import gather_env_vector
        return gather_env_vector(self.engine.stats[STAT_KEYS.index("return")])

    def close(self):
        if not self._is_closed:
            self.engine.close()
        super().close()


class PIDRateController(Transform):
    """Stand-in for the reference's PIDrate action transform
    (omni_drones/utils/torchrl/transforms.py:404-459).  The arithmetic of that transform and of
    the rate PID it wraps (lee_position_controller.py:476-550) is fused into hs_tick_kernel;
    this object only switches the base env to raw-action mode and re-declares the action spec
    the way the reference's ``transform_input_spec`` does."""

    def __init__(self, controller=None, action_key=("agents", "action")):
        super().__init__([], in_keys_inv=[("info", "drone_state")])
        self.controller, self.action_key = controller, action_key

    def set_parent(self, env):
        super().set_parent(env)
        base = env.base_env if hasattr(env, "base_env") else env
        base.action_is_raw = True

    def transform_input_spec(self, input_spec):
        spec = input_spec[("_action_spec", *self.action_key)]
        input_spec[("_action_spec", *self.action_key)] = UnboundedContinuousTensorSpec(
            tuple(spec.shape[:-1]) + (4,), device=spec.device)
        return input_spec
