"""HideAndSeek_envgen: the same tick + the Adaptive Environment Generator control plane.

Reference: omni_drones/envs/hide_and_seek/hideandseek_envgen.py (GenBuffer :209-377, class :379,
reset :875-1013, generator update inside the reward :1241-1246, :1302-1333).  The per-tick
arithmetic is identical to HideAndSeek (the reference file is a copy; its smoothness reward is
not gated by use_deployment, :1284) and runs in the same kernels with the variant flags of
hs_config.  What this module adds is the host-side control plane, which in the reference is
numpy on the host as well:

  * reset draws tasks from {uniform sampling} U {perturbed samples of the archive}, re-drawn
    every `eval_iter` episodes, uniform tasks on a prefix of the env index space;
  * on the done tick the per-env success becomes a weight; every `eval_iter` episodes the tasks
    whose mean weight lies in [R_min, R_max] enter the archive (cap 5000, farthest-point
    sampling -- restated in torch, the reference calls dgl.geometry.farthest_point_sampler).

Like the reference (hideandseek_envgen.py:896-898) this variant assumes that all envs reset
together; a partial `_reset` mask raises.
"""
import math
import os
from typing import Optional

import numpy as np
import torch

from ..compat import CompositeSpec, TensorDict, UnboundedContinuousTensorSpec
from .hideandseek import STAT_KEYS, HideAndSeek


def farthest_point_sampling(points: torch.Tensor, k: int, start: int = 0) -> torch.Tensor:
    """Indices of k points chosen greedily to maximise the minimum pairwise distance (the
    standard FPS recurrence; replaces dgl.geometry.farthest_point_sampler, which also starts
    from a fixed first point)."""
    n = points.shape[0]
    k = min(k, n)
    idx = torch.empty(k, dtype=torch.long, device=points.device)
    dist = torch.full((n,), float("inf"), device=points.device, dtype=points.dtype)
    cur = start
    for i in range(k):
        idx[i] = cur
        d = ((points - points[cur]) ** 2).sum(-1)
        dist = torch.minimum(dist, d)
        cur = int(torch.argmax(dist))
    return idx


class GenBuffer:
    """Task archive (hideandseek_envgen.py:209-377).  A task = [drone xyz * A, evader xyz, cylinder xyz * C]."""

    def __init__(self, num_agents: int, num_cylinders: int, arena_size=0.9, cylinder_size=0.1, max_height=1.2,
                 buffer_length: int = 5000, rng: Optional[np.random.Generator] = None):
        self.num_agents, self.num_cylinders = num_agents, num_cylinders
        self.task_dim = 3 * num_agents + 3 + 3 * num_cylinders     # the reference hard-codes 18 + 3A (C = 5)
        self._history_buffer = np.zeros((0, self.task_dim), dtype=np.float32)
        self._state_buffer = np.zeros((0, self.task_dim), dtype=np.float32)
        self._weight_buffer = np.zeros((0, 1), dtype=np.float32)
        self._temp_state_buffer, self._temp_weight_buffer = [], []
        self.buffer_length, self.eps = buffer_length, 1e-5
        self.arena_size, self.cylinder_size, self.max_height = arena_size, cylinder_size, max_height
        self.grid_size = 2 * cylinder_size
        self.num_grid = int(arena_size * 2 / self.grid_size)
        self.rng = rng or np.random.default_rng()
        n, half = self.num_grid, self.num_grid // 2
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        self.blocked = np.sqrt((ii - half) ** 2 + (jj - half) ** 2) >= half     # cells outside the arena circle

    # -- archive bookkeeping ---------------------------------------------------------------
    def insert(self, states):
        self._temp_state_buffer.extend(np.array(states, copy=True))

    def insert_weights(self, weights: torch.Tensor):
        self._temp_weight_buffer.append(weights.detach().float().cpu().numpy().reshape(-1, 1))

    def update(self):
        self._state_buffer = np.array(self._temp_state_buffer)
        self._weight_buffer = np.stack(self._temp_weight_buffer, axis=-1).mean(-1)
        self._temp_state_buffer, self._temp_weight_buffer = [], []

    def insert_history(self, states: np.ndarray):
        if len(states) == 0:
            return
        all_states = np.concatenate([self._history_buffer, states.astype(np.float32)])
        if all_states.shape[0] > self.buffer_length:
            lo, hi = all_states.min(0), all_states.max(0)
            normed = torch.from_numpy((all_states - lo) / (hi - lo + self.eps))
            keep = farthest_point_sampling(normed, self.buffer_length).numpy()
            all_states = all_states[keep]
        self._history_buffer = all_states

    def save_task(self, model_dir, episode):
        np.save(os.path.join(model_dir, f"history_{episode}.npy"), self._history_buffer)

    # -- sampling --------------------------------------------------------------------------
    def _cells(self, xy):
        return np.clip(np.round(xy / self.grid_size).astype(int) + self.num_grid // 2, 0, self.num_grid - 1)

    def _valid(self, task) -> bool:
        """sanity_check (:185-207): every object must claim its own free grid cell."""
        A = self.num_agents
        xy = np.concatenate([task[:3 * A].reshape(-1, 3)[:, :2], task[3 * A:3 * A + 3].reshape(-1, 3)[:, :2],
                             task[3 * A + 3:].reshape(-1, 3)[:, :2]])
        c = self._cells(xy)
        occ = self.blocked.copy()
        before = occ.sum()
        occ[c[:, 0], c[:, 1]] = True
        return occ.sum() - before >= len(c)

    def sample(self, num_tasks):
        return self._history_buffer[self.rng.integers(0, self._history_buffer.shape[0], num_tasks)]

    def samplenearby(self, num_tasks, expand_cylinders, expand_step):
        A, C = self.num_agents, self.num_cylinders
        origin = self.sample(num_tasks)
        cb = int(self.arena_size / self.grid_size) * self.grid_size
        bxy = self.arena_size / math.sqrt(2.0) - 0.1
        drone_b = [[-bxy, bxy], [-bxy, bxy], [self.max_height - 0.1, self.max_height + 0.1]]   # sic: z in [1.1, 1.3]
        cyl_b = [[-cb, cb], [-cb, cb], [-20.0, self.max_height / 2]]
        bounds = np.array(drone_b * (A + 1) + cyl_b * C)
        out = []
        for i in range(num_tasks):
            for _ in range(10):
                noise_dt = self.rng.uniform(-1, 1, size=3 * A + 3) * expand_step
                noise_c = np.zeros(3 * C)
                if expand_cylinders:
                    nc = np.zeros((C, 3))
                    nc[:, :2] = self.rng.choice([-1, 0, 1], size=(C, 2)) * self.grid_size
                    noise_c = nc.reshape(-1)
                cand = np.clip(origin[i] + np.concatenate([noise_dt, noise_c]), bounds[:, 0], bounds[:, 1])
                if self._valid(cand):
                    out.append(cand)
                    break
        out = np.array(out, dtype=np.float32).reshape(-1, self.task_dim)
        if 0 < out.shape[0] < num_tasks:
            extra = out[self.rng.integers(0, out.shape[0], num_tasks - out.shape[0])]
            out = np.concatenate([extra, out])
        elif out.shape[0] == 0:
            out = origin.astype(np.float32)
        return out


class GenBufferDevice:
    """The same archive with everything resident on the GPU: tasks, weights
    and the archive are device tensors; the two loops that are host work in the reference run as
    kernels - `samplenearby` (hideandseek_envgen.py:322-372, a Python loop over tasks with up to
    ten numpy retries each) is hs_gen_sample_nearby, the archive cap (:300-314,
    dgl.geometry.farthest_point_sampler) is hs_fps.  Same interface as GenBuffer; arrays are
    torch tensors instead of numpy arrays."""

    def __init__(self, num_agents: int, num_cylinders: int, arena_size=0.9, cylinder_size=0.1, max_height=1.2,
                 buffer_length: int = 5000, seed: int = 0, device="cuda:0", task_offset: int = 0):
        import ctypes as C
        from .. import _lib
        self._C, self._lib = C, _lib
        self.device = torch.device(device)
        self.num_agents, self.num_cylinders = num_agents, num_cylinders
        self.task_dim = 3 * num_agents + 3 + 3 * num_cylinders
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self._history_buffer, self._state_buffer, self._weight_buffer = z(0, self.task_dim), z(0, self.task_dim), z(0, 1)
        self._temp_state_buffer, self._temp_weight_buffer = [], []
        self.buffer_length, self.eps = buffer_length, 1e-5
        self.params = _lib.hs_gen_params()
        self.params.num_agents, self.params.num_cylinders = num_agents, num_cylinders
        self.params.arena_size, self.params.grid_size, self.params.max_height = arena_size, 2 * cylinder_size, max_height
        self.params.num_grid = int(arena_size * 2 / (2 * cylinder_size))
        self.params.seed = int(seed) & (2 ** 64 - 1)
        self.params.task_offset = int(task_offset)      # sharded jobs: rank r draws tasks [offset, offset + n)
        self.epoch = 0
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed))

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- archive bookkeeping ---------------------------------------------------------------
    def insert(self, states: torch.Tensor):
        self._temp_state_buffer.append(states.detach().to(self.device, torch.float32).clone())

    def insert_weights(self, weights: torch.Tensor):
        self._temp_weight_buffer.append(weights.detach().float().reshape(-1, 1).clone())

    def update(self):
        self._state_buffer = torch.cat(self._temp_state_buffer)
        self._weight_buffer = torch.stack(self._temp_weight_buffer, dim=-1).mean(-1)
        self._temp_state_buffer, self._temp_weight_buffer = [], []

    def fps(self, points: torch.Tensor, k: int, start: int = 0) -> torch.Tensor:
        pts = points.to(self.device, torch.float32).contiguous()
        n, dim = pts.shape
        k = min(k, n)
        idx = torch.empty(k, dtype=torch.int32, device=self.device)
        scratch = torch.empty(int(self._lib.lib.hs_fps_scratch_bytes(n)), dtype=torch.uint8, device=self.device)
        self._lib.check(self._lib.lib.hs_fps(pts.data_ptr(), n, dim, k, start, idx.data_ptr(), scratch.data_ptr(),
                                             self._stream()), "hs_fps")
        self._keep = (pts, scratch)
        return idx.long()

    def insert_history(self, states: torch.Tensor):
        if states.shape[0] == 0:
            return
        all_states = torch.cat([self._history_buffer, states.to(self.device, torch.float32)])
        if all_states.shape[0] > self.buffer_length:
            lo, hi = all_states.min(0).values, all_states.max(0).values
            normed = (all_states - lo) / (hi - lo + self.eps)
            all_states = all_states[self.fps(normed, self.buffer_length)]
        self._history_buffer = all_states

    def save_task(self, model_dir, episode):
        np.save(os.path.join(model_dir, f"history_{episode}.npy"), self._history_buffer.cpu().numpy())

    # -- sampling --------------------------------------------------------------------------
    def sample(self, num_tasks):
        idx = torch.randint(0, self._history_buffer.shape[0], (num_tasks,), device=self.device, generator=self.gen)
        return self._history_buffer[idx]

    def samplenearby(self, num_tasks, expand_cylinders, expand_step, return_valid: bool = False):
        C = self._C
        self.epoch += 1
        self.params.expand_cylinders, self.params.expand_step = int(bool(expand_cylinders)), float(expand_step)
        hist = self._history_buffer.contiguous()
        out = torch.empty(num_tasks, self.task_dim, dtype=torch.float32, device=self.device)
        valid = torch.empty(num_tasks, dtype=torch.uint8, device=self.device)
        self._lib.check(self._lib.lib.hs_gen_sample_nearby(C.byref(self.params), hist.data_ptr(), hist.shape[0], num_tasks,
                                                           C.c_uint64(self.epoch), out.data_ptr(), valid.data_ptr(),
                                                           self._stream()), "hs_gen_sample_nearby")
        self._keep2 = hist
        if return_valid:
            return out, valid.bool()
        ok = valid.bool()
        # rows whose ten attempts all failed are re-drawn from the accepted ones (:361-366); no host sync:
        # a random accepted row per slot, used only where needed
        n_ok = ok.sum()
        order = torch.argsort((~ok).to(torch.uint8), stable=True)                  # accepted rows first
        pick = (torch.rand(num_tasks, device=self.device, generator=self.gen) * n_ok.clamp(min=1)).long()
        fallback = torch.where(n_ok > 0, out[order[pick]], hist[0].expand_as(out))
        return torch.where(ok.unsqueeze(-1), out, fallback)


class HideAndSeek_envgen(HideAndSeek):
    VARIANT_ENVGEN = True
    EXTRA_STATS = ("success_buffer", "success_unif", "history_buffer", "add_history", "ratio_unif")

    def _design_scene(self):
        super()._design_scene()
        t = self.cfg.task
        self.use_particle_generator = bool(t.use_particle_generator)
        self.ratio_unif, self.eval_iter = float(t.ratio_unif), int(t.eval_iter)
        self.success_threshold = float(t.success_threshold)
        self.expand_cylinders, self.expand_step = bool(t.expand_cylinders), float(t.expand_step)
        self.R_min, self.R_max = float(t.R_min), float(t.R_max)
        self.update_iter = 0
        self.num_unif = self.num_envs
        seed = int(self.cfg.seed or 0)
        # env.device_generator=1 (default): archive, perturbation sampler and FPS on the GPU
        self.device_generator = True if self.cfg.env.device_generator is None else bool(self.cfg.env.device_generator)
        if self.device_generator:
            self.gen_buffer = GenBufferDevice(self.num_agents, self.num_cylinders, t.arena_size, t.cylinder.size,
                                              t.max_height, seed=seed, device=self.device,
                                              task_offset=int(self.cfg.env.env_offset or 0))
        else:
            self.gen_buffer = GenBuffer(self.num_agents, self.num_cylinders, t.arena_size, t.cylinder.size, t.max_height,
                                        rng=np.random.default_rng(seed))
        self.all_tasks = None
        self._host_progress = 0
        # sharded job: this rank owns the envs [env_offset, env_offset + E) of `global_num_envs` (default: a single shard).
        # The reference's uniform / archive split is a PREFIX of the global env order (hideandseek_envgen.py:1241-1246).
        self.env_offset = int(self.cfg.env.env_offset or 0)
        self.global_num_envs = int(self.cfg.env.global_num_envs or (self.env_offset + self.num_envs))

    def _stat_keys(self):
        extra = list(self.EXTRA_STATS)
        for i in range(self.num_cylinders + 1):
            extra += [f"ratio_cylinders_{i}", f"success_cylinders_{i}"]
        return tuple(STAT_KEYS) + tuple(extra)

    def _set_specs(self):
        super()._set_specs()
        E, dev = self.num_envs, self.device
        for k in self._stat_keys()[len(STAT_KEYS):]:
            self.stats.set(k, torch.zeros(E, 1, device=dev))

    # ------------------------------------------------------------------ reset
    def _sample_reset(self, n: int):
        base = super()._sample_reset(n)              # uniform sampling (:860-873) + orientations
        if not (self.use_random_cylinder and self.use_particle_generator):
            return base
        A, dev = self.num_agents, self.device
        if self.update_iter == 0:
            # global split first, then this shard's part of it (single shard: identical to the reference)
            Eg = max(self.global_num_envs, self.env_offset + n)
            num_buffer_g = min(self.gen_buffer._history_buffer.shape[0], int(Eg * (1 - self.ratio_unif)))
            self.num_unif = int(min(max(Eg - num_buffer_g - self.env_offset, 0), n))
            num_buffer = n - self.num_unif
            unif = torch.cat([base["drone_pos"][:self.num_unif].reshape(self.num_unif, -1),
                              base["target_pos"][:self.num_unif].reshape(self.num_unif, -1),
                              base["cyl_pos"][:self.num_unif].reshape(self.num_unif, -1)], dim=-1)
            if self.device_generator:                          # everything stays on the GPU
                if num_buffer > 0:
                    near = self.gen_buffer.samplenearby(num_buffer, self.expand_cylinders, self.expand_step)
                    self.all_tasks = torch.cat([unif, near])
                else:
                    self.all_tasks = unif.clone()
            else:
                unif = unif.cpu().numpy()
                if num_buffer > 0:
                    near = self.gen_buffer.samplenearby(num_buffer, self.expand_cylinders, self.expand_step)
                    self.all_tasks = np.concatenate([unif, near])
                else:
                    self.all_tasks = unif
            self.gen_buffer.insert(self.all_tasks)
        if self.device_generator:
            tasks = self.all_tasks
        else:
            tasks = torch.from_numpy(np.ascontiguousarray(self.all_tasks)).to(dev).float()
        base["drone_pos"] = tasks[:, :3 * A].reshape(n, A, 3)
        base["target_pos"] = tasks[:, 3 * A:3 * A + 3].reshape(n, 3)
        base["cyl_pos"] = tasks[:, 3 * A + 3:].reshape(n, -1, 3)
        base["active_cylinders"] = (base["cyl_pos"][..., 2] > 0.0).float().sum(-1, keepdim=True)
        return base

    def _reset(self, tensordict=None, init=None, **kwargs):
        if tensordict is not None and "_reset" in tensordict and not bool(tensordict.get("_reset").all()):
            raise RuntimeError("HideAndSeek_envgen resets all envs together (hideandseek_envgen.py:896-898)")
        self._host_progress = 0
        # `self.stats[env_ids] = 0.` (hideandseek_envgen.py:997) also clears the generator's stats: add_history and the
        # per-cylinder-count ratios are non-zero only from the update tick to the next reset
        for k in self._stat_keys()[len(STAT_KEYS):]:
            self.stats[k].zero_()
        return super()._reset(None if tensordict is None else tensordict.exclude("_reset"), init=init, **kwargs)

    # ------------------------------------------------------------------ step
    def _step(self, tensordict):
        out = super()._step(tensordict)
        self._host_progress += 1
        success = self.engine.stats[0]
        st = self.stats
        done_tick = self._host_progress >= self.max_episode_length  # == torch.any(done), without a host sync
        sharded = self.global_num_envs > self.num_envs
        if sharded and done_tick:
            # the values the scripts harvest (EpisodeStats reads the post-done step) are means over the GLOBAL prefix /
            # suffix: one all_reduce of four numbers per episode; between episode ends the per-rank means stand in
            from ..parallel import global_sum
            nu = self.num_unif
            acc = global_sum(torch.stack([success[:nu].sum(), success.new_tensor(float(nu)),
                                          success[nu:].sum(), success.new_tensor(float(self.num_envs - nu))]))
            st["success_unif"].fill_(0).add_(acc[0] / acc[1].clamp(min=1))
            st["success_buffer"].fill_(0).add_(acc[2] / acc[3].clamp(min=1))
        elif self.num_unif < self.num_envs:
            st["success_buffer"].fill_(0).add_(success[self.num_unif:].mean())
            if self.num_unif > 0:
                st["success_unif"].fill_(0).add_(success[:self.num_unif].mean())
        elif not sharded:
            st["success_buffer"].zero_()
            st["success_unif"].copy_(success.unsqueeze(-1))
        else:
            st["success_buffer"].zero_()
            st["success_unif"].fill_(0).add_(success.mean())
        if done_tick:
            self._on_episode_end(success)
        st["history_buffer"].fill_(float(len(self.gen_buffer._history_buffer)))
        st["ratio_unif"].fill_(self.ratio_unif)
        return out

    def _on_episode_end(self, success: torch.Tensor):
        """hideandseek_envgen.py:1302-1330."""
        from ..parallel import global_mean
        if float(global_mean(success)) > self.success_threshold:
            self.ratio_unif = 1.0
        self.gen_buffer.insert_weights(success)
        self.update_iter += 1
        if self.update_iter < self.eval_iter:
            return
        self.update_iter = 0
        self.gen_buffer.update()
        w = self.gen_buffer._weight_buffer.reshape(-1)
        if self.device_generator:
            active = self.active_cylinders.reshape(-1)
            wd = w.to(active.device)
            for i in range(self.num_cylinders + 1):            # device-side means, no host sync
                sel = (active == i).float()
                cnt = sel.sum()
                self.stats[f"ratio_cylinders_{i}"].fill_(0).add_(cnt / sel.numel())
                self.stats[f"success_cylinders_{i}"].fill_(0).add_((wd * sel).sum() / cnt.clamp(min=1))
        else:
            active = self.active_cylinders.reshape(-1).cpu().numpy()
            for i in range(self.num_cylinders + 1):
                sel = active == i
                self.stats[f"ratio_cylinders_{i}"].fill_(float(sel.mean()))
                self.stats[f"success_cylinders_{i}"].fill_(float(w[sel].mean()) if sel.any() else 0.0)
        keep = (w <= self.R_max) & (w >= self.R_min)
        kept = self.gen_buffer._state_buffer[keep]
        if self.device_generator:
            # sharded job: the archive is REPLICATED - every rank inserts the tasks all ranks evaluated, in
            # rank order, and the (deterministic) farthest point sampling keeps the copies identical
            from ..parallel import gather_rows
            kept = gather_rows(kept)
        self.gen_buffer.insert_history(kept)
        self.stats["add_history"].fill_(0).add_(torch.as_tensor(keep.sum(), device=self.stats["add_history"].device))
