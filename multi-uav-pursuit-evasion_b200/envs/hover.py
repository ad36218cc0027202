"""Hover: single-Crazyflie plumbing task (BASELINE.json configs[0]).

Reference: omni_drones/envs/single/hover.py:40-523.  The vehicle side of the tick (PIDrate
transform, rate PID, rotor model, rigid-body step) is the SAME fused kernel as HideAndSeek,
instantiated with one pursuer and no cylinders; the evader slot idles far away.  The
Hover-specific observation (20 wide), reward and the 39 logging stats are ONE small post-kernel
(hs_hover_post, csrc/hs_hover.cuh) on what the tick left in the engine's buffers.
Not built: payload / mass randomisation, observation latency and noise options (all off in
cfg/task/Hover.yaml)."""
import ctypes as C

import torch

from .. import _lib
from ..compat import CompositeSpec, TensorDict, UnboundedContinuousTensorSpec
from ..config import build_hs_config, load_drone_params
from ..engine import HsEngine
from .agent_spec import AgentSpec
from .hideandseek import DroneView
from .isaac_env import IsaacEnv


class Hover(IsaacEnv):
    # declaration order of the reference's stats spec (hover.py:239-279) = slot order of hs_hover_post
    STAT_KEYS = ("return", "pos_bonus", "head_bonus", "reward_pos", "reward_up", "reward_vel", "reward_acc", "reward_jerk",
                 "episode_len", "pos_error", "heading_alignment", "uprightness", "action_smoothness",
                 "linear_v_max", "angular_v_max", "linear_a_max", "angular_a_max", "linear_jerk_max", "angular_jerk_max",
                 "linear_v_mean", "angular_v_mean", "linear_a_mean", "angular_a_mean", "linear_jerk_mean", "angular_jerk_mean",
                 "motor1", "motor2", "motor3", "motor4", "cmd_r", "cmd_p", "cmd_y", "cmd_thrust",
                 "target_r_rate", "target_p_rate", "target_y_rate", "real_r_rate", "real_p_rate", "real_y_rate")
    # MultirotorBase.intrinsics_spec (multirotor.py:78-88): all zeros unless a randomisation is configured (:652-697)
    INTRINSICS = (("mass", 1), ("inertia", 3), ("KF", 4), ("KM", 4), ("tau_up", 4), ("tau_down", 4), ("drag_coef", 1),
                  ("rotor_offset", 1))

    def _design_scene(self):
        t = self.cfg.task
        for flag in ("add_noise", "latency", "action_noise"):
            if t[flag]:
                raise NotImplementedError(f"Hover option {flag}=true is not built (off in cfg/task/Hover.yaml)")
        self.reward_distance_scale = t.reward_distance_scale
        self.reward_v_scale, self.reward_acc_scale, self.reward_jerk_scale = t.reward_v_scale, t.reward_acc_scale, t.reward_jerk_scale
        self.linear_vel_max, self.linear_acc_max = t.linear_vel_max, t.linear_acc_max
        self.time_encoding = bool(t.time_encoding)
        self.time_encoding_dim = 4 if self.time_encoding else 0
        self.use_omega, self.use_motor = bool(t.omega), bool(t.motor)
        params = load_drone_params()
        # Hover keeps PhysX's default velocity limits (no v_drone clamp): robots/config.py:36-38
        self._hs_cfg = build_hs_config(self.num_envs, num_agents=1, num_cylinders=0, obs_max_cylinder=0,
                                       use_tp_net=False, max_episode_length=self.max_episode_length, dt=self.dt,
                                       max_linear_velocity=1000.0, drone_params=params)
        self.engine = HsEngine(self._hs_cfg, self.device, num_output_sets=2)
        self.drone = DroneView(self.engine, params, 1, self.device)
        self.action_is_raw = False
        E, dev = self.num_envs, self.device
        self.target_pos = torch.tensor([[0.0, 0.0, 1.0]], device=dev)
        self.target_heading = torch.zeros(E, 1, 3, device=dev)         # quat_axis(target_rot, 0) at reset (hover.py:317-319)
        self.alpha = 0.8
        self._far = torch.tensor([50.0, 50.0, 0.5], device=dev).expand(E, 3).contiguous()   # idle evader slot
        self.obs_dim = 16 + (3 if self.use_omega else 0) + (4 if self.use_motor else 0) + self.time_encoding_dim
        # post-kernel buffers: stats [39, E], last values + episode sums [12, E], throttle difference of the tick [E, 1]
        self._stats = torch.zeros(_lib.HS_HOVER_NUM_STATS, E, device=dev)
        self._state = torch.zeros(_lib.HS_HOVER_NUM_STATE, E, device=dev)
        self._throttle_diff = torch.zeros(E, 1, device=dev)
        for b in self.engine._bufs:
            b.throttle_diff = self._throttle_diff.data_ptr()
        self.engine._bind(self.engine.cur)
        self._obs_buf = [torch.zeros(E, 1, self.obs_dim, device=dev) for _ in range(2)]
        self._reward_buf = [torch.zeros(E, 1, 1, device=dev) for _ in range(2)]
        self._done_buf = [torch.zeros(E, 1, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._flip = 0
        hp = _lib.hs_hover_params()
        hp.reward_distance_scale, hp.reward_v_scale = self.reward_distance_scale, self.reward_v_scale
        hp.reward_acc_scale, hp.reward_jerk_scale = self.reward_acc_scale, self.reward_jerk_scale
        hp.linear_vel_max, hp.linear_acc_max, hp.alpha = self.linear_vel_max, self.linear_acc_max, self.alpha
        hp.target_pos[:] = [0.0, 0.0, 1.0]
        hp.time_encoding, hp.omega, hp.motor = int(self.time_encoding), int(self.use_omega), int(self.use_motor)
        self._hp = hp

    def _set_specs(self):
        E, dev = self.num_envs, self.device
        U = UnboundedContinuousTensorSpec
        intr_spec = CompositeSpec({k: U((1, w), device=dev) for k, w in self.INTRINSICS})
        self.observation_spec = CompositeSpec({"agents": CompositeSpec({
            "observation": U((1, self.obs_dim), device=dev), "intrinsics": intr_spec})}).expand(E).to(dev)
        self.action_spec = CompositeSpec({"agents": CompositeSpec({"action": self.drone.action_spec.unsqueeze(0)})}).expand(E).to(dev)
        self.reward_spec = CompositeSpec({"agents": CompositeSpec({"reward": U((1, 1))})}).expand(E).to(dev)
        self.agent_spec["drone"] = AgentSpec("drone", 1, observation_key=("agents", "observation"),
                                             action_key=("agents", "action"), reward_key=("agents", "reward"))
        stats_spec = CompositeSpec({k: U(1) for k in self.STAT_KEYS}).expand(E).to(dev)
        info_spec = CompositeSpec({"drone_state": U((1, 13), device=dev),
                                   "prev_action": self.drone.action_spec.unsqueeze(0)}).expand(E).to(dev)
        self.observation_spec["stats"] = stats_spec
        self.observation_spec["info"] = info_spec
        # live views of the post-kernel's stats rows (the reference hands out live references too)
        self.stats = TensorDict({k: self._stats[i].unsqueeze(-1) for i, k in enumerate(self.STAT_KEYS)}, [E], dev)
        self.intrinsics = TensorDict({k: torch.zeros(E, 1, w, device=dev) for k, w in self.INTRINSICS}, [E], dev)
        self.info = TensorDict({"drone_state": self.engine.out["drone_state"], "prev_action": self.engine.prev_action}, [E], dev)

    @property
    def progress_buf(self):
        return self.engine.get_state(_lib.FIELD_PROGRESS)

    def _post(self, out, with_reward: bool):
        """hs_hover_post on the engine's latest outputs; returns (observation, reward, done) of this tick."""
        self._flip ^= 1
        obs, rew, done = self._obs_buf[self._flip], self._reward_buf[self._flip], self._done_buf[self._flip]
        io = _lib.hs_hover_io()
        io.observation, io.reward, io.done = obs.data_ptr(), rew.data_ptr(), done.data_ptr()
        io.stats, io.state, io.target_heading = self._stats.data_ptr(), self._state.data_ptr(), self.target_heading.data_ptr()
        self._hp.with_reward = 1 if with_reward else 0
        _lib.check(_lib.lib.hs_hover_post(self.engine._h, C.byref(self._hp), C.byref(io), self.engine._stream()), "hs_hover_post")
        self.info.set("drone_state", out["drone_state"])
        return obs, rew, done.view(torch.bool)

    def _obs_td(self, obs):
        return TensorDict({"agents": {"observation": obs, "intrinsics": self.intrinsics}, "stats": self.stats, "info": self.info},
                          self.batch_size, self.device)

    def _reset(self, tensordict=None, init=None, **kwargs):
        E, dev = self.num_envs, self.device
        mask = tensordict.get("_reset").reshape(E) if tensordict is not None and "_reset" in tensordict else None
        last_stats = self.stats.clone()
        if init is None:
            lo, hi = torch.tensor([-1.0, -1.0, 0.05], device=dev), torch.tensor([1.0, 1.0, 2.0], device=dev)
            pos = lo + (hi - lo) * torch.rand(E, 1, 3, device=dev)
            rlo = torch.tensor([-0.2, -0.2, 0.0], device=dev) * torch.pi
            rhi = torch.tensor([0.2, 0.2, 0.5], device=dev) * torch.pi
            rpy = rlo + (rhi - rlo) * torch.rand(E, 1, 3, device=dev)
            r, p, y = rpy.unbind(-1)
            cy, sy, cp, sp, cr, sr = torch.cos(y / 2), torch.sin(y / 2), torch.cos(p / 2), torch.sin(p / 2), torch.cos(r / 2), torch.sin(r / 2)
            rot = torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                               cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], dim=-1)
            init = dict(drone_pos=pos, drone_rot=rot)
        m = slice(None) if mask is None else mask
        # hover.py:296-332: stats and the last_* values of the reset envs are zeroed (init_vels = 0); the episode sums are
        # re-created for ALL envs (`self.linear_v_episode = torch.zeros_like(...)`, a reference quirk); target heading =
        # quat_axis(euler(0, 0, 0), 0) = x
        self._stats[:, m] = 0.0
        self._state[:6, m] = 0.0
        self._state[6:] = 0.0
        self.target_heading[m] = torch.tensor([1.0, 0.0, 0.0], device=dev)
        out = self.engine.reset(mask, init["drone_pos"], init["drone_rot"], self._far, torch.zeros(E, 0, 3, device=dev))
        obs, _, _ = self._post(out, with_reward=False)
        td = TensorDict({}, self.batch_size, dev)
        td.update(self._obs_td(obs))
        td.set("stats", last_stats)
        td.set("truncated", out["truncated"])
        return td

    def _step(self, tensordict):
        eng = self.engine
        action = tensordict.get(("agents", "action")).reshape(self.num_envs, 1, 4).contiguous()
        if self.action_is_raw:
            out = eng.step_pre(action, raw=True, reset_pid=tensordict.get("done", None))
            tensordict.set(("agents", "action"), out["rotor_cmds"])
            tensordict.set("ctbr", out["ctbr"])
            tensordict.set("target_rate", out["target_rate"])
            tensordict.set(("info", "prev_action"), eng.prev_action)
        else:
            # rotor commands applied directly: the keys the PIDrate transform would have written are taken from the input
            # tensordict when present (hover.py:349-359 reads td['ctbr'] / td['target_rate'])
            nxt_set = eng.sets[eng.next_index()]
            nxt_set["rotor_cmds"].copy_(action)
            for k, w in (("ctbr", 4), ("target_rate", 3)):
                v = tensordict.get(k, None)
                nxt_set[k].copy_(v.reshape(self.num_envs, 1, w)) if v is not None else nxt_set[k].zero_()
            out = eng.step_pre(action, raw=False, reset_pid=None)
        obs, rew, done = self._post(out, with_reward=True)
        nxt = self._obs_td(obs)
        nxt.set(("agents", "reward"), rew)
        nxt.set("done", done)
        return TensorDict({"next": nxt}, self.batch_size, self.device)

    def close(self):
        if not self._is_closed:
            self.engine.close()
        super().close()
