"""Hover: single-Crazyflie plumbing task (BASELINE.json configs[0]).

Reference: omni_drones/envs/single/hover.py:40-523.  The vehicle side of the tick (PIDrate
transform, rate PID, rotor model, rigid-body step) is the SAME fused kernel as HideAndSeek,
instantiated with one pursuer and no cylinders; the evader slot idles far away.  The
Hover-specific observation (20 wide) and reward are a handful of elementwise torch ops on the
kernel's `drone_state` output -- this task is plumbing (64..100 envs), not the measured path.
Not built: payload / mass randomisation, observation latency and noise options (all off in
cfg/task/Hover.yaml)."""
import torch

from .. import _lib
from ..compat import CompositeSpec, TensorDict, UnboundedContinuousTensorSpec
from ..config import build_hs_config, load_drone_params
from ..engine import HsEngine
from .agent_spec import AgentSpec
from .hideandseek import DroneView
from .isaac_env import IsaacEnv


def _quat_rotate(q, v):
    w, u = q[..., 0:1], q[..., 1:4]
    return v * (2.0 * w ** 2 - 1.0) + torch.linalg.cross(u, v, dim=-1) * w * 2.0 + u * (u * v).sum(-1, keepdim=True) * 2.0


class Hover(IsaacEnv):
    def _design_scene(self):
        t = self.cfg.task
        for flag in ("omega", "motor", "add_noise", "latency", "action_noise"):
            if t[flag]:
                raise NotImplementedError(f"Hover option {flag}=true is not built (off in cfg/task/Hover.yaml)")
        self.reward_distance_scale = t.reward_distance_scale
        self.reward_v_scale, self.reward_acc_scale, self.reward_jerk_scale = t.reward_v_scale, t.reward_acc_scale, t.reward_jerk_scale
        self.linear_vel_max, self.linear_acc_max = t.linear_vel_max, t.linear_acc_max
        self.time_encoding = bool(t.time_encoding)
        self.time_encoding_dim = 4 if self.time_encoding else 0
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
        self.target_heading = torch.zeros(E, 1, 3, device=dev)
        self.target_heading[..., 0] = 1.0
        self.alpha = 0.8
        self._far = torch.tensor([50.0, 50.0, 0.5], device=dev).expand(E, 3).contiguous()   # idle evader slot
        self.last_linear_v = torch.zeros(E, 1, device=dev)
        self.last_linear_a = torch.zeros(E, 1, device=dev)

    STAT_KEYS = ("return", "pos_bonus", "head_bonus", "reward_pos", "reward_vel", "reward_acc", "reward_jerk",
                 "pos_error", "heading_alignment", "uprightness", "action_smoothness", "episode_len",
                 "linear_v_max", "linear_a_max", "linear_jerk_max")

    def _set_specs(self):
        E, dev = self.num_envs, self.device
        U = UnboundedContinuousTensorSpec
        obs_dim = 3 + 7 + 6 + self.time_encoding_dim
        self.observation_spec = CompositeSpec({"agents": CompositeSpec({"observation": U((1, obs_dim), device=dev)})}).expand(E).to(dev)
        self.action_spec = CompositeSpec({"agents": CompositeSpec({"action": self.drone.action_spec.unsqueeze(0)})}).expand(E).to(dev)
        self.reward_spec = CompositeSpec({"agents": CompositeSpec({"reward": U((1, 1))})}).expand(E).to(dev)
        self.agent_spec["drone"] = AgentSpec("drone", 1, observation_key=("agents", "observation"),
                                             action_key=("agents", "action"), reward_key=("agents", "reward"))
        stats_spec = CompositeSpec({k: U(1) for k in self.STAT_KEYS}).expand(E).to(dev)
        info_spec = CompositeSpec({"drone_state": U((1, 13), device=dev),
                                   "prev_action": self.drone.action_spec.unsqueeze(0)}).expand(E).to(dev)
        self.observation_spec["stats"] = stats_spec
        self.observation_spec["info"] = info_spec
        self.stats = stats_spec.zero()
        self.info = TensorDict({"drone_state": self.engine.out["drone_state"], "prev_action": self.engine.prev_action}, [E], dev)

    @property
    def progress_buf(self):
        return self.engine.get_state(_lib.FIELD_PROGRESS)

    def _obs(self, out):
        ds = out["drone_state"]                                   # [E,1,13]
        pos, quat, linvel = ds[..., :3], ds[..., 3:7], ds[..., 7:10]
        ex = torch.zeros_like(pos); ex[..., 0] = 1.0
        ez = torch.zeros_like(pos); ez[..., 2] = 1.0
        self.heading, self.up = _quat_rotate(quat, ex), _quat_rotate(quat, ez)
        self.rpos = self.target_pos - pos
        self.rheading = self.target_heading - self.heading
        parts = [self.rpos, quat, linvel, self.heading, self.up]
        prog = self.progress_buf
        if self.time_encoding:
            parts.append((prog / self.max_episode_length).reshape(-1, 1, 1).expand(-1, 1, 4))
        self.linear_v = torch.linalg.vector_norm(linvel, dim=-1)
        self.linear_a = torch.abs(self.linear_v - self.last_linear_v) / self.dt
        self.linear_jerk = torch.abs(self.linear_a - self.last_linear_a) / self.dt
        for k, v in (("linear_v_max", self.linear_v), ("linear_a_max", self.linear_a), ("linear_jerk_max", self.linear_jerk)):
            self.stats[k].copy_(torch.max(self.stats[k], v))
        self.last_linear_v, self.last_linear_a = self.linear_v.clone(), self.linear_a.clone()
        self.info.set("drone_state", ds)
        self._progress = prog
        return TensorDict({"agents": {"observation": torch.cat(parts, dim=-1)}, "stats": self.stats, "info": self.info},
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
        for v in self.stats.values():
            v[m] = 0.0
        self.last_linear_v[m] = 0.0
        self.last_linear_a[m] = 0.0
        out = self.engine.reset(mask, init["drone_pos"], init["drone_rot"], self._far, torch.zeros(E, 0, 3, device=dev))
        td = TensorDict({}, self.batch_size, dev)
        td.update(self._obs(out))
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
            out = eng.step_pre(action, raw=False, reset_pid=None)
        nxt = self._obs(out)
        # reward, hover.py:439-523
        pos_error = torch.linalg.vector_norm(self.rpos, dim=-1)
        head_error = torch.linalg.vector_norm(self.rheading, dim=-1)
        reward_pos = -pos_error * self.reward_distance_scale
        bonus = ((pos_error <= 0.02) * 10).float()
        reward_head = -head_error * (bonus > 0)
        head_bonus = ((head_error <= 0.02) * 10 * (bonus > 0)).float()
        reward_up = torch.square((self.up[..., 2] + 1) / 2)
        reward_v = self.reward_v_scale * (bonus > 0) * (self.linear_v < self.linear_vel_max)
        reward_acc = self.reward_acc_scale * (bonus > 0) * (self.linear_a < self.linear_acc_max)
        reward_jerk = self.reward_jerk_scale * (bonus > 0) * (-self.linear_jerk)
        reward = reward_pos + bonus + reward_head + head_bonus + reward_up + reward_v + reward_acc + reward_jerk
        done = (self._progress >= self.max_episode_length).unsqueeze(-1)
        st = self.stats
        st["pos_error"].lerp_(pos_error, 1 - self.alpha)
        st["heading_alignment"].lerp_((self.heading * self.target_heading).sum(-1), 1 - self.alpha)
        st["uprightness"].lerp_(self.up[..., 2], 1 - self.alpha)
        st["return"].add_(reward)
        for k, v in (("reward_pos", reward_pos), ("pos_bonus", bonus), ("head_bonus", head_bonus),
                     ("reward_vel", reward_v), ("reward_acc", reward_acc), ("reward_jerk", reward_jerk)):
            st[k].copy_(v.float() if torch.is_tensor(v) else torch.full_like(st[k], float(v)))
        st["episode_len"].copy_(self._progress.unsqueeze(1))
        nxt.set(("agents", "reward"), reward.unsqueeze(-1))
        nxt.set("done", done)
        return TensorDict({"next": nxt}, self.batch_size, self.device)

    def close(self):
        if not self._is_closed:
            self.engine.close()
        super().close()
