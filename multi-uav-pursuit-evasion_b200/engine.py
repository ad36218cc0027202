"""Device-memory plumbing around the C ABI: allocates the arena and the reference-facing
output tensors with PyTorch, binds them, and launches the kernels on torch's current stream.

PyTorch is used here for memory, streams and (elsewhere) torch.distributed only; all
arithmetic of the environment tick happens in libhs_b200.so.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import lib, check, hs_buffers, hs_config

_ALIGN_WORDS = 32            # 128 B: keeps every output tensor TMA-bulk-store aligned


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class OutputSet:
    """One set of reference-facing tensors carved out of a single float32 slab (+ a byte slab)."""

    @staticmethod
    def shapes(cfg: hs_config) -> Dict[str, tuple]:
        E, A = cfg.num_envs, cfg.num_agents
        K, F, H = cfg.obs_max_cylinder, cfg.future_step, cfg.history_step
        D = 20 + (3 * F if cfg.use_tp_net else 0)
        FD = 7 + 3 * A + (3 * cfg.num_cylinders if cfg.use_obstacles else 0)
        # the tensors a policy consumes come first, so that they form one contiguous prefix of
        # the slab (a single D2H copy moves observation + reward when the policy lives on the host)
        shapes = {
            "state_self": (E, A, 1, D), "state_others": (E, A, max(A - 1, 0), 3),
            "obs_cylinders": (E, A, K, 5), "reward": (E, A, 1),
            "state_drones": (E, A, D), "drone_state": (E, A, 13), "rotor_cmds": (E, A, 4),
            "ctbr": (E, A, 4), "target_rate": (E, A, 3), "action_error": (E, A),
        }
        if cfg.use_tp_net:
            shapes.update({"tp_input": (E, H, FD), "tp_groundtruth": (E, 3)})
        return shapes

    def __init__(self, cfg: hs_config, device, views: Optional[Dict[str, torch.Tensor]] = None):
        E = cfg.num_envs
        shapes = self.shapes(cfg)
        if views is not None:
            # slot of a RolloutStorage: every tensor is one [E, ...] row of a time-major [T, E, ...] tensor
            self.slab, self.policy_words, self.result_words = None, 0, 0
            self.t = dict(views)
            return
        offs, total = {}, 0
        for k, s in shapes.items():
            n = 1
            for d in s:
                n *= d
            offs[k] = (total, n)
            total += (n + _ALIGN_WORDS - 1) // _ALIGN_WORDS * _ALIGN_WORDS
            if k == "reward":
                self.policy_words = total      # slab[:policy_words] = observation + reward
        self.slab = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.t: Dict[str, torch.Tensor] = {}
        for k, s in shapes.items():
            o, n = offs[k]
            self.t[k] = self.slab[o:o + n].view(*s)
        self.bytes = torch.zeros(3, E, 1, dtype=torch.uint8, device=device)
        self.t["done"] = self.bytes[0].view(torch.bool)
        self.t["tp_done"] = self.bytes[1].view(torch.bool)
        self.t["truncated"] = self.bytes[2].view(torch.bool)
        self.result_words = total

    def __getitem__(self, k):
        return self.t[k]


class RolloutStorage:
    """Time-major rollout buffers the tick kernels write into directly: one
    ``[T + 1, E, ...]`` tensor per output; tick ``t`` of a rollout binds row ``t`` as its output
    set, so after T ticks ``[:T]`` IS the rollout - nothing is cloned or stacked.  Row ``T`` is the
    scratch set that receives the outputs of resets (the reference stores the pre-reset ``next``
    in the rollout and feeds the post-reset observation to the following step).  Every row starts
    128-byte aligned (TMA bulk stores)."""

    def __init__(self, cfg: hs_config, device, num_steps: int):
        self.T = int(num_steps)
        E = cfg.num_envs
        self.data: Dict[str, torch.Tensor] = {}
        for k, shp in OutputSet.shapes(cfg).items():
            n = 1
            for d in shp:
                n *= d
            pitch = (n + _ALIGN_WORDS - 1) // _ALIGN_WORDS * _ALIGN_WORDS
            flat = torch.zeros(self.T + 1, max(pitch, 1), dtype=torch.float32, device=device)
            self.data[k] = flat[:, :n].view(self.T + 1, *shp)
        for k in ("done", "tp_done", "truncated"):
            self.data[k] = torch.zeros(self.T + 1, E, 1, dtype=torch.uint8, device=device).view(torch.bool)
        # input side of tick t when a FusedPolicy is attached to the engine: what the actor / critic produced from
        # the observation the tick started from (rows of the same time-major rollout)
        A = cfg.num_agents
        self.policy: Dict[str, torch.Tensor] = {
            "action": torch.zeros(self.T + 1, E, A, 4, dtype=torch.float32, device=device),
            "logp": torch.zeros(self.T + 1, E, A, 1, dtype=torch.float32, device=device),
            "action_mean": torch.zeros(self.T + 1, E, A, 4, dtype=torch.float32, device=device),
            "state_value": torch.zeros(self.T + 1, E, A, 1, dtype=torch.float32, device=device)}

    def slot(self, cfg: hs_config, device, i: int) -> OutputSet:
        return OutputSet(cfg, device, {k: v[i] for k, v in self.data.items()})

    def batch(self) -> Dict[str, torch.Tensor]:
        """``[E, T, ...]`` views (env-major indexing, time-major memory) of the finished rollout."""
        return {k: v[:self.T].transpose(0, 1) for k, v in self.data.items()}

    def policy_batch(self) -> Dict[str, torch.Tensor]:
        """``[E, T, ...]`` views of the attached policy's outputs (action, logp, action_mean, state_value)."""
        return {k: v[:self.T].transpose(0, 1) for k, v in self.policy.items()}


class HsEngine:
    """Owns the buffers of one environment batch on one GPU and drives the kernels."""

    def __init__(self, cfg: hs_config, device="cuda:0", num_output_sets: int = 2, rollout_steps: Optional[int] = None):
        device = torch.device(device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.HsError("HsEngine needs a CUDA device: the environment step has no CPU path "
                               f"(requested {device}, torch.cuda.is_available()={torch.cuda.is_available()})")
        self.cfg = cfg
        self.device = device
        self.E, self.A, self.C = cfg.num_envs, cfg.num_agents, cfg.num_cylinders
        with torch.cuda.device(device):
            h = C.c_void_p()
            check(lib.hs_create(C.byref(cfg), C.byref(h)), "hs_create")
        self._h = h
        n = lib.hs_arena_floats(C.byref(cfg))
        self.arena = torch.zeros(n, dtype=torch.float32, device=device)
        self.stats = torch.zeros(_lib.HS_NUM_STATS, self.E, dtype=torch.float32, device=device)
        self.prev_action = torch.zeros(self.E, self.A, 4, dtype=torch.float32, device=device)
        self.v_prey = torch.full((1,), 1.3, dtype=torch.float32, device=device)
        # device scalar read by every tick (hs_buffers.smoothness_coef): the host refreshes it when update_epoch changes
        self.smoothness_coef = torch.full((1,), float(cfg.smoothness_coef), dtype=torch.float32, device=device)
        self.storage: Optional[RolloutStorage] = None
        if rollout_steps:
            # rollout mode: set t = row t of the time-major rollout tensors, set T = reset scratch
            self.storage = RolloutStorage(cfg, device, int(rollout_steps))
            self.sets = [self.storage.slot(cfg, device, i) for i in range(self.storage.T + 1)]
            self._slot = -1                    # rollout row written by the latest tick
        else:
            self.sets = [OutputSet(cfg, device) for _ in range(max(1, num_output_sets))]
        self._bufs = [self._make_bufs(i) for i in range(len(self.sets))]
        self.cur = len(self.sets) - 1          # index of the set holding the latest outputs
        # Host-side upper bound of the env progress counters (None = unknown): `done` is progress >= max_episode_length
        # (hideandseek.py:1008-1010), so callers can rule out "some env is done" without reading the device.
        self.host_max_progress: Optional[int] = 0
        self._keep = []                        # keeps caller tensors alive across async launches
        # identity quaternions so that an un-reset arena is still well formed
        ident = torch.zeros(self.E, self.A, 4, device=device)
        ident[..., 0] = 1.0
        self._bind(self.cur)
        self.set_state(_lib.FIELD_DRONE_ROT, ident)

    # ------------------------------------------------------------------ plumbing
    def _make_bufs(self, i: int) -> hs_buffers:
        s = self.sets[i]
        prev = self.sets[(i - 1) % len(self.sets)]
        b = hs_buffers()
        b.arena, b.stats = _ptr(self.arena), _ptr(self.stats)
        for k in ("state_self", "obs_cylinders", "state_drones", "reward", "drone_state",
                  "rotor_cmds", "ctbr", "target_rate", "action_error"):
            setattr(b, k, _ptr(s[k]))
        b.state_others = _ptr(s["state_others"]) if self.A > 1 else None
        if self.cfg.use_tp_net:
            b.tp_input, b.tp_input_prev = _ptr(s["tp_input"]), _ptr(prev["tp_input"])
            b.tp_groundtruth = _ptr(s["tp_groundtruth"])
        b.tp_done, b.done, b.truncated = _ptr(s["tp_done"]), _ptr(s["done"]), _ptr(s["truncated"])
        b.prev_action, b.v_prey = _ptr(self.prev_action), _ptr(self.v_prey)
        b.smoothness_coef = _ptr(self.smoothness_coef)
        if getattr(self, "tp_ring", None) is not None:
            b.tp_ring, b.tp_ring_pos = _ptr(self.tp_ring), _ptr(self.tp_ring_pos)
        return b

    def set_tp_ring(self, on: bool = True):
        """Ring form of the TP window (hs_buffers.tp_ring, include/hs_b200.h): the tick writes the new frame twice into
        ``tp_ring`` [E, 2H, FD] instead of shifting a chronological [E,H,FD] tensor (1088 B/env/tick less HBM traffic for
        the reference's shape).  The window is then :meth:`tp_window`, a strided view valid until the next tick, and
        ``out["tp_input"]`` is NOT written.  Lane-per-env tick mapping only (num_agents >= 3); call before the first
        reset / tick."""
        if not self.cfg.use_tp_net:
            raise _lib.HsError("set_tp_ring: config has use_tp_net == 0")
        if on:
            H, FD = self.cfg.history_step, self.sets[0]["tp_input"].shape[-1]
            self.tp_ring = torch.zeros(self.E, 2 * H, FD, dtype=torch.float32, device=self.device)
            self.tp_ring_pos = torch.zeros((self.E + 31) // 32, dtype=torch.int32, device=self.device)
        else:
            self.tp_ring = self.tp_ring_pos = None
        self._bufs = [self._make_bufs(i) for i in range(len(self.sets))]
        self._graphs = None
        self._bind(self.cur)
        return self

    def tp_window(self) -> torch.Tensor:
        """The chronological TP window [E,H,FD] of the latest tick: ``out["tp_input"]`` in plain mode, the strided view
        into the ring in ring mode (reads the ring position from the device: one 4-byte D2H)."""
        if getattr(self, "tp_ring", None) is None:
            return self.out["tp_input"]
        p, H = int(self.tp_ring_pos[0].item()), self.cfg.history_step
        return self.tp_ring[:, p:p + H, :]

    def _bind(self, i: int, prev: Optional[int] = None):
        """Binds output set i; the previous TP window is read from set ``prev`` (default: the set
        that holds the latest outputs)."""
        b = self._bufs[i]
        if self.cfg.use_tp_net:
            b.tp_input_prev = _ptr(self.sets[self.cur if prev is None else prev]["tp_input"])
        check(lib.hs_bind_buffers(self._h, C.byref(b)), "hs_bind_buffers")
        self.cur = i

    def next_index(self, reset: bool = False) -> int:
        """Index of the set the next tick (or reset) will write."""
        if self.storage is None:
            return (self.cur + 1) % len(self.sets)
        return self.storage.T if reset else (self._slot + 1) % self.storage.T

    def maybe_done(self) -> bool:
        """False: certainly no env reports done after the latest tick (host-side counter, no device sync)."""
        return self.host_max_progress is None or self.host_max_progress >= self.cfg.max_episode_length

    def refresh_host_progress(self) -> int:
        """Re-reads the largest progress counter from the device (one sync; after partial resets / state injection)."""
        self.host_max_progress = int(self.get_state(_lib.FIELD_PROGRESS).max().item())
        return self.host_max_progress

    def _advance(self, reset: bool = False):
        if not reset and self.host_max_progress is not None:
            self.host_max_progress += 1
        i = self.next_index(reset)
        if self.storage is not None and not reset:
            self._slot = i
        self._bind(i)

    # ------------------------------------------------------------------ policy next to the tick
    def attach_policy(self, actor, critic=None, deterministic: bool = False, defer_critic: bool = False):
        """Makes ``actor`` (a :class:`~mupe_b200.policy.FusedPolicy`) - and optionally ``critic`` - part of the tick:
        :meth:`policy_tick` and the graphs captured afterwards run actor -> critic -> hs_step_pre -> hs_step_post_tp
        on the observation the previous tick produced, i.e. one whole rollout step (``policy(td)`` + ``env.step(td)``
        in the reference's collector) without a host round trip.  The actor's noise is drawn in its kernel
        (``actor.seed(...)``); ``deterministic`` takes the mode instead.  Outputs land in ``policy_out`` (per output
        set; rows of the rollout storage in rollout mode)."""
        if actor.head_dim != 4 or not actor.is_actor:
            raise _lib.HsError("attach_policy: the actor must be a DiagGaussian head over the 4 CTBR commands")
        E, A, dev = self.E, self.A, self.device
        self._actor, self._critic, self._deterministic = actor, critic, bool(deterministic)
        self.parallel_critic = True         # graphs: critic as a parallel branch beside actor -> tick
        # defer_critic (rollout mode): the per-step launches leave the critic out; :class:`PolicyRolloutGraph` evaluates
        # it ONCE per rollout over the T stored observations (the critic has no state: same values, one large launch)
        self._defer_critic = bool(defer_critic) and critic is not None
        if self._defer_critic and self.storage is None:
            raise _lib.HsError("attach_policy(defer_critic=True) needs rollout mode (rollout_steps=T)")
        if not deterministic and getattr(actor, "rng_state", None) is None:
            actor.seed(0)
        if self.storage is not None:
            pol = self.storage.policy
            self.policy_out = [{k: v[i] for k, v in pol.items()} for i in range(len(self.sets))]
        else:
            z = lambda w: torch.zeros(E, A, w, dtype=torch.float32, device=dev)
            self.policy_out = [dict(action=z(4), logp=z(1), action_mean=z(4), state_value=z(1)) for _ in self.sets]
        self._graphs = None
        return self

    def _launch_policy(self, prev: int, i: int, fork: bool = False):
        """actor (+ critic) on the observation held by set ``prev``; results into policy_out[i].  ``fork`` (graph capture):
        the critic, which nothing in the tick depends on, goes to a side stream and becomes a parallel branch of the
        graph; returns the event the caller joins on after the tick."""
        obs, po = self.sets[prev], self.policy_out[i]
        others = obs["state_others"] if self.A > 1 else None
        joined = None
        if getattr(self, "_defer_critic", False):
            self._actor.forward(obs["state_self"], others, obs["obs_cylinders"], sample=not self._deterministic,
                                out={"head": po["action_mean"], "action": po["action"], "logp": po["logp"]})
            return None
        if self._critic is not None and fork:
            cur = torch.cuda.current_stream(self.device)
            if getattr(self, "_critic_stream", None) is None:
                self._critic_stream = torch.cuda.Stream(self.device)
            start = torch.cuda.Event()
            start.record(cur)
            self._critic_stream.wait_event(start)
            with torch.cuda.stream(self._critic_stream):
                self._critic.forward(obs["state_self"], others, obs["obs_cylinders"], out={"head": po["state_value"]})
                joined = torch.cuda.Event()
                joined.record(self._critic_stream)
        self._actor.forward(obs["state_self"], others, obs["obs_cylinders"], sample=not self._deterministic,
                            out={"head": po["action_mean"], "action": po["action"], "logp": po["logp"]})
        if self._critic is not None and not fork:
            self._critic.forward(obs["state_self"], others, obs["obs_cylinders"], out={"head": po["state_value"]})
        return joined

    def policy_tick(self, tp_weights=None, reset_pid: Optional[torch.Tensor] = None) -> OutputSet:
        """One rollout step without CUDA graphs: actor -> critic -> tick -> fused predictor (4 launches)."""
        prev, i = self.cur, self.next_index()
        self._launch_policy(prev, i)
        self._policy_launches = getattr(self, "_policy_launches", 0) + (2 if self._critic is not None else 1)
        if self.cfg.use_tp_net:
            if tp_weights is None:
                raise _lib.HsError("policy_tick: use_tp_net needs tp_weights (fused predictor)")
            return self.step_fused(self.policy_out[i]["action"], tp_weights, raw=True, reset_pid=reset_pid)
        return self.step_pre(self.policy_out[i]["action"], raw=True, reset_pid=reset_pid)

    @property
    def rollout_slot(self) -> int:
        """Rollout row the latest tick wrote (rollout mode); the rollout is complete at T - 1."""
        return self._slot

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def out(self) -> OutputSet:
        return self.sets[self.cur]

    # ------------------------------------------------------------------ hot path
    def step_pre(self, action: torch.Tensor, raw: bool = True, reset_pid: Optional[torch.Tensor] = None) -> OutputSet:
        E, A = self.E, self.A
        if action.shape != (E, A, 4) or action.dtype != torch.float32 or not action.is_contiguous() \
                or action.device != self.device:
            raise _lib.HsError(f"action must be a contiguous float32 [{E},{A},4] tensor on {self.device}")
        rp = None
        if reset_pid is not None:
            rp = reset_pid.reshape(E)
            if rp.dtype == torch.bool:
                rp = rp.view(torch.uint8)
            if rp.dtype != torch.uint8 or not rp.is_contiguous():
                rp = rp.to(torch.uint8).contiguous()
        self._advance()
        self._keep = [action, rp]
        check(lib.hs_step_pre(self._h, action.data_ptr(), 1 if raw else 0, _ptr(rp), self._stream()), "hs_step_pre")
        return self.out

    def wait_host(self):
        """Completes a ``step_host(..., sync=False)``: the host views it returned are valid afterwards."""
        st = getattr(self, "_host_stream", None)
        if st is not None:
            check(lib.hs_host_io_wait(self._h, st), "hs_host_io_wait")
            self._host_stream = None

    def step_host(self, action_host: torch.Tensor, weights=None, raw: bool = True, reset_pid: Optional[torch.Tensor] = None,
                  sync: bool = True):
        """hs_step_host_io: one C-ABI call from HOST buffers to HOST buffers - pinned action in, H2D,
        tick (+ fused predictor), D2H of observation / reward / done, stream sync.  Returns
        (host mirror of the policy-facing slab prefix [state_self | state_others | obs_cylinders |
        reward] as a dict of views, done).  The mirror is pinned and reused by the next call."""
        E, A = self.E, self.A
        if self.storage is not None:
            raise _lib.HsError("step_host: the host-buffer tick needs slab output sets (rollout_steps=None)")
        if getattr(self, "_host", None) is None:
            npol = self.sets[0].policy_words
            mirror = torch.empty(npol, dtype=torch.float32).pin_memory()
            self._host = dict(mirror=mirror, done=torch.empty(E, dtype=torch.uint8).pin_memory(),
                              staging=torch.empty(E, A, 4, dtype=torch.float32, device=self.device))
            self._host["staging_ptr"] = self._host["staging"].data_ptr()
        hm = self._host
        self._advance()
        cached = hm.setdefault("io", {}).get(self.cur)
        if cached is None:
            # per output set: the io struct and the host views never change (static buffers)
            out = self.out
            base = out.slab.data_ptr()
            io = _lib.hs_host_io()
            views = {}
            for k in ("state_self", "state_others", "obs_cylinders", "reward"):
                t = out.t[k]
                off = (t.data_ptr() - base) // 4
                views[k] = hm["mirror"][off:off + t.numel()].view(t.shape)
                setattr(io, k, views[k].data_ptr() if t.numel() else None)
            io.done = hm["done"].data_ptr()
            cached = hm["io"][self.cur] = (io, C.byref(io), views)
        io, io_ref, views = cached
        io.action = action_host.data_ptr()
        rp = None
        if reset_pid is not None:
            rp = reset_pid.reshape(E)
            rp = rp.view(torch.uint8) if rp.dtype == torch.bool else rp.to(torch.uint8)
            rp = rp.contiguous()
        self._keep = [action_host, rp]
        st = self._stream()
        fn = lib.hs_step_host_io if sync else lib.hs_step_host_io_async
        rc = fn(self._h, io_ref, 1 if raw else 0, _ptr(rp), C.byref(weights) if weights is not None else None,
                hm["staging_ptr"], st)
        if rc != 0:
            check(rc, "hs_step_host_io")
        self._host_stream = None if sync else st      # sync=False: call wait_host() before reading the views
        return views, hm["done"]

    def step_fused(self, action: torch.Tensor, weights: "_lib.hs_tp_weights", raw: bool = True,
                   reset_pid: Optional[torch.Tensor] = None, pred_out: Optional[torch.Tensor] = None) -> OutputSet:
        """hs_step_fused: tick + fused predictor in one call - ONE kernel launch for batches of at most one
        32-env tile per SM (hs_tick_tp_fused_kernel), otherwise hs_step_pre + hs_step_post_tp."""
        E, A = self.E, self.A
        if action.shape != (E, A, 4) or action.dtype != torch.float32 or not action.is_contiguous() \
                or action.device != self.device:
            raise _lib.HsError(f"action must be a contiguous float32 [{E},{A},4] tensor on {self.device}")
        rp = None
        if reset_pid is not None:
            rp = reset_pid.reshape(E)
            rp = rp.view(torch.uint8) if rp.dtype == torch.bool else rp.to(torch.uint8)
            rp = rp.contiguous()
        if pred_out is not None:
            assert pred_out.shape == (E, 3 * self.cfg.future_step) and pred_out.is_contiguous()
        self._advance()
        self._keep = [action, rp, pred_out]
        check(lib.hs_step_fused(self._h, action.data_ptr(), 1 if raw else 0, _ptr(rp), C.byref(weights), _ptr(pred_out),
                                self._stream()), "hs_step_fused")
        return self.out

    def rollout_fused(self, actions: torch.Tensor, num_ticks: int, weights: "_lib.hs_tp_weights", raw: bool = True,
                      pred_out: Optional[torch.Tensor] = None) -> OutputSet:
        """hs_rollout_fused: ``num_ticks`` control ticks (tick + predictor) in ONE kernel launch for actions that are
        already on the device - ``actions`` [T,E,A,4], or [E,A,4] applied every tick.  Tick t writes the next output set
        (rollout mode: the next row of the time-major storage), exactly like ``num_ticks`` calls of :meth:`step_fused`;
        returns the set of the last tick.  ``pred_out`` [T,E,3F] optional."""
        E, A, T = self.E, self.A, int(num_ticks)
        per_tick = actions.dim() == 4
        ok = actions.shape == ((T, E, A, 4) if per_tick else (E, A, 4))
        if not ok or actions.dtype != torch.float32 or not actions.is_contiguous() or actions.device != self.device:
            raise _lib.HsError(f"actions must be a contiguous float32 [{T},{E},{A},4] or [{E},{A},4] tensor on {self.device}")
        F3 = 3 * self.cfg.future_step
        if pred_out is not None:
            assert pred_out.shape == (T, E, F3) and pred_out.is_contiguous() and pred_out.dtype == torch.float32
        n = len(self.sets) if self.storage is None else self.storage.T
        if getattr(self, "_sets_table", None) is None or self._sets_table_src is not self._bufs:
            raw_bytes = b"".join(bytes(self._bufs[i]) for i in range(n))
            self._sets_table = torch.frombuffer(bytearray(raw_bytes), dtype=torch.uint8).to(self.device)
            self._sets_table_src = self._bufs
        first = self.next_index()
        prev_tp = self.sets[self.cur]["tp_input"]
        self._keep = [actions, pred_out]
        check(lib.hs_rollout_fused(self._h, self._sets_table.data_ptr(), n, first, prev_tp.data_ptr(), actions.data_ptr(),
                                   E * A * 4 if per_tick else 0, 1 if raw else 0, T, C.byref(weights), _ptr(pred_out), E * F3,
                                   self._stream()), "hs_rollout_fused")
        last = (first + T - 1) % n
        if self.host_max_progress is not None:
            self.host_max_progress += T
        if self.storage is not None:
            self._slot = last
        self._bind(last, prev=(last - 1) % n)
        return self.out

    def step_post(self, tp_pred: torch.Tensor) -> OutputSet:
        F3 = 3 * self.cfg.future_step
        tp_pred = tp_pred.reshape(self.E, F3)
        if tp_pred.dtype != torch.float32 or not tp_pred.is_contiguous():
            tp_pred = tp_pred.float().contiguous()
        self._keep.append(tp_pred)
        check(lib.hs_step_post(self._h, tp_pred.data_ptr(), self._stream()), "hs_step_post")
        return self.out

    def tp_weights(self, module) -> Optional["_lib.hs_tp_weights"]:
        """hs_tp_weights for a TP_net-shaped module (``lstm``: 1-layer LSTM hidden 64, ``fc``:
        Linear), or None when the module does not have that shape.  Pointers are taken from the
        live parameters, so in-place optimiser updates are seen by the next tick."""
        lstm, fc = getattr(module, "lstm", None), getattr(module, "fc", None)
        if not isinstance(lstm, torch.nn.LSTM) or not isinstance(fc, torch.nn.Linear):
            return None
        if self.A > 3:              # the fused predictor kernels cover up to 3 pursuers: module forward + hs_step_post
            return None
        if lstm.num_layers != 1 or lstm.bidirectional or lstm.hidden_size != 64 or not lstm.batch_first \
                or lstm.proj_size != 0 or not lstm.bias:
            return None
        ps = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, fc.weight, fc.bias]
        if any(p.dtype != torch.float32 or p.device != self.device or not p.is_contiguous() for p in ps):
            return None
        w = _lib.hs_tp_weights()
        w.weight_ih, w.weight_hh, w.bias_ih, w.bias_hh, w.fc_weight, w.fc_bias = [p.data_ptr() for p in ps]
        w.input_size, w.hidden_size, w.output_size = lstm.input_size, lstm.hidden_size, fc.out_features
        if w.input_size != 7 + 3 * self.A or w.output_size != 3 * self.cfg.future_step or self.cfg.use_obstacles:
            return None
        return w

    def step_post_tp(self, weights: "_lib.hs_tp_weights", pred_out: Optional[torch.Tensor] = None) -> OutputSet:
        """Second half with the predictor fused into the kernel (no cuDNN call)."""
        if pred_out is not None:
            assert pred_out.shape == (self.E, 3 * self.cfg.future_step) and pred_out.is_contiguous()
            self._keep.append(pred_out)
        check(lib.hs_step_post_tp(self._h, C.byref(weights), _ptr(pred_out), self._stream()), "hs_step_post_tp")
        return self.out

    def set_predictor_variant(self, variant: int):
        """-1: auto (default); 0: fp32 FFMA kernel; 1: 3xTF32 mma.sync kernel; 2: 3xTF32 tcgen05/TMEM kernel
        (128-env tiles); 3: 3xTF32 tcgen05 kernel with the gates on M and 32-env tiles (small batches);
        4: as 3 with two tiles ping-ponging per CTA (larger batches); 5: the 32-env tile as two ping-ponging 16-env
        halves (small-batch default, and the predictor half of the one-launch tick)."""
        check(lib.hs_set_option(self._h, _lib.HS_OPT_PREDICTOR_VARIANT, int(variant)), "hs_set_option")
        self._graphs = None             # captured graphs hold the old kernel

    def set_tick_mapping(self, mapping: int):
        """HS_OPT_TICK_MAPPING: 0 auto (4 lanes per env below 32768 envs, one lane per env above and for more than 3
        pursuers), 1 always 4 lanes per env, 2 always one lane per env (hs_tick_wide_kernel).  Same results bit for bit."""
        check(lib.hs_set_option(self._h, _lib.HS_OPT_TICK_MAPPING, int(mapping)), "hs_set_option")
        self._graphs = None

    def set_rollout_variant(self, variant: int):
        """HS_OPT_ROLLOUT_VARIANT: 0 auto, else the number of ticks the predictor warps advance per pass (1: hs_rollout_fused_kernel;
        2, 3: hs_rollout_pair_kernel)."""
        check(lib.hs_set_option(self._h, _lib.HS_OPT_ROLLOUT_VARIANT, int(variant)), "hs_set_option")

    def set_exact_math(self, on: bool):
        """HS_OPT_EXACT_MATH: run the tick with the IEEE-arithmetic build of the kernel (parity evidence; ~2x slower)."""
        check(lib.hs_set_option(self._h, _lib.HS_OPT_EXACT_MATH, 1 if on else 0), "hs_set_option")
        self._graphs = None

    # ------------------------------------------------------------------ CUDA graphs
    def capture_tick_graphs(self, tp_weights=None, raw: bool = True):
        """Captures one CUDA graph per output set holding a whole tick (hs_step_pre and, with the
        predictor, hs_step_post_tp).  Inputs are read from the static buffers ``graph_action``
        [E,A,4] and ``graph_reset_pid`` [E] uint8; replay with :meth:`replay_tick`.  Must be
        called after the first reset (the history-initialising first frame is not captured)."""
        if self.cfg.use_tp_net and tp_weights is None:
            raise _lib.HsError("capture_tick_graphs: use_tp_net needs tp_weights (fused predictor)")
        dev = self.device
        self.graph_action = torch.zeros(self.E, self.A, 4, dtype=torch.float32, device=dev)
        self.graph_reset_pid = torch.zeros(self.E, dtype=torch.uint8, device=dev)
        self._graph_weights = tp_weights
        self._graph_raw = raw
        self._graphs = {}
        self._graph_kernels = (2 if self.cfg.use_tp_net else 1)
        self._graph_policy_kernels = 0
        if getattr(self, "_actor", None) is not None:
            self._graph_policy_kernels = 2 if self._critic is not None else 1
        self._graph_replays = getattr(self, "_graph_replays", 0)      # cumulative, like hs_launch_count
        if self.storage is None:
            for i in range(len(self.sets)):
                self._capture((i - 1) % len(self.sets), i)
        return self

    def _capture(self, prev: int, i: int):
        """One CUDA graph for 'tick into set i with the previous TP window in set prev'."""
        dev = self.device
        keep = self.cur
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        self._bind(i, prev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            st = torch.cuda.current_stream(dev).cuda_stream
            action = self.graph_action
            joined = None
            if getattr(self, "_actor", None) is not None:
                joined = self._launch_policy(prev, i, fork=self.parallel_critic)
                action = self.policy_out[i]["action"]
            n0 = int(lib.hs_launch_count(self._h))
            if self.cfg.use_tp_net:
                # one launch (hs_tick_tp_fused_kernel) when the batch qualifies, else tick + predictor
                check(lib.hs_step_fused(self._h, action.data_ptr(), 1 if self._graph_raw else 0,
                                        self.graph_reset_pid.data_ptr(), C.byref(self._graph_weights), None, st),
                      "hs_step_fused (capture)")
            else:
                check(lib.hs_step_pre(self._h, action.data_ptr(), 1 if self._graph_raw else 0,
                                      self.graph_reset_pid.data_ptr(), st), "hs_step_pre (capture)")
            self._graph_kernels = int(lib.hs_launch_count(self._h)) - n0
            if joined is not None:
                torch.cuda.current_stream(dev).wait_event(joined)
        self._graphs[(prev, i)] = g
        self._graph_captures = getattr(self, "_graph_captures", 0) + 1
        self._bind(keep, keep)
        return g

    def replay_tick(self) -> OutputSet:
        """One tick from ``graph_action`` / ``graph_reset_pid`` with a single graph launch.  Graphs are
        keyed by (set holding the previous TP window, set written); in rollout mode the pairs are
        captured on first use (t-1 -> t, T-1 -> 0 and reset scratch -> t)."""
        prev, i = self.cur, self.next_index()
        g = self._graphs.get((prev, i))
        if g is None:
            g = self._capture(prev, i)
        self._advance()
        g.replay()
        self._graph_replays += 1
        return self.out

    def reset(self, mask: Optional[torch.Tensor], drone_pos, drone_rot, target_pos, cyl_pos) -> OutputSet:
        E, A, Cc = self.E, self.A, self.C
        f = lambda t, shape: t.to(self.device, torch.float32).reshape(shape).contiguous()
        drone_pos, drone_rot = f(drone_pos, (E, A, 3)), f(drone_rot, (E, A, 4))
        target_pos = f(target_pos, (E, 3))
        cyl_pos = f(cyl_pos, (E, Cc, 3)) if Cc > 0 else None
        m = None
        if mask is not None:
            m = mask.to(self.device).reshape(E)
            m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
            m = m.contiguous()
        self._advance(reset=True)
        self.host_max_progress = 0 if m is None else None       # a partial reset leaves the other envs' counters unknown
        self._keep = [m, drone_pos, drone_rot, target_pos, cyl_pos]
        check(lib.hs_reset(self._h, _ptr(m), drone_pos.data_ptr(), drone_rot.data_ptr(), target_pos.data_ptr(),
                           _ptr(cyl_pos), self._stream()), "hs_reset")
        return self.out

    def sample_reset(self, dist: "_lib.hs_reset_dist", epoch: int) -> Dict[str, torch.Tensor]:
        """Device-side reset sampler (hs_sample_reset): one launch that draws the initial poses
        of all envs of a random-cylinder episode - what HideAndSeek._reset_idx samples with torch
        distributions and a per-env host randperm loop (hideandseek.py:609-697, 106-119).
        Returns the dict HsEngine.reset / HideAndSeek._reset(init=...) consume; the buffers are
        reused by the next call."""
        E, A, Cc = self.E, self.A, self.C
        if getattr(self, "_rs_bufs", None) is None:
            z = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=self.device)
            self._rs_bufs = dict(drone_pos=z(E, A, 3), drone_rot=z(E, A, 4), target_pos=z(E, 3),
                                 cyl_pos=z(E, max(Cc, 1), 3)[:, :Cc], active_cylinders=z(E, 1))
        b = self._rs_bufs
        check(lib.hs_sample_reset(self._h, C.byref(dist), C.c_uint64(int(epoch) & (2 ** 64 - 1)), b["drone_pos"].data_ptr(),
                                  b["drone_rot"].data_ptr(), b["target_pos"].data_ptr(),
                                  _ptr(b["cyl_pos"]) if Cc > 0 else None, b["active_cylinders"].data_ptr(),
                                  self._stream()), "hs_sample_reset")
        return b

    # ------------------------------------------------------------------ state views
    _SHAPES = {
        _lib.FIELD_DRONE_POS: lambda s: (s.E, s.A, 3), _lib.FIELD_DRONE_ROT: lambda s: (s.E, s.A, 4),
        _lib.FIELD_DRONE_LINVEL: lambda s: (s.E, s.A, 3), _lib.FIELD_DRONE_ANGVEL: lambda s: (s.E, s.A, 3),
        _lib.FIELD_THROTTLE: lambda s: (s.E, s.A, 4), _lib.FIELD_PID_INTEG: lambda s: (s.E, s.A, 3),
        _lib.FIELD_PID_LAST_RATE: lambda s: (s.E, s.A, 3), _lib.FIELD_TARGET_POS: lambda s: (s.E, 3),
        _lib.FIELD_TARGET_VEL: lambda s: (s.E, 3), _lib.FIELD_CYL_POS: lambda s: (s.E, s.C, 3),
        _lib.FIELD_PROGRESS: lambda s: (s.E,),
    }

    def get_state(self, field: int) -> torch.Tensor:
        out = torch.empty(self._SHAPES[field](self), dtype=torch.float32, device=self.device)
        if out.numel():
            check(lib.hs_state_get(self._h, field, out.data_ptr(), self._stream()), "hs_state_get")
        return out

    def set_state(self, field: int, value: torch.Tensor):
        v = value.to(self.device, torch.float32).reshape(self._SHAPES[field](self)).contiguous()
        if field == _lib.FIELD_PROGRESS:
            self.host_max_progress = None
        if v.numel():
            check(lib.hs_state_set(self._h, field, v.data_ptr(), self._stream()), "hs_state_set")
            self._keep.append(v)

    @property
    def launches(self) -> int:
        """Kernels of libhs_b200.so launched so far (graph replays counted per captured kernel)."""
        n = int(lib.hs_launch_count(self._h))
        n += getattr(self, "_policy_launches", 0) + getattr(self, "_graph_replays", 0) * getattr(self, "_graph_policy_kernels", 0)
        n += getattr(self, "_uncounted", 0)                  # RotatingRolloutGraph replays
        if getattr(self, "_graph_captures", 0):
            n += (self._graph_replays - getattr(self, "_graph_captures", 0)) * self._graph_kernels   # capture calls counted once each
        return n

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            lib.hs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RotatingRolloutGraph:
    """ONE CUDA graph holding ``ticks`` consecutive control ticks that rotate over several engines (independent env
    batches on one GPU): tick t runs on ``engines[t % len(engines)]``.  With the actions resident on the device
    (``graph_action`` of every engine, or an attached policy) nothing in a rollout needs the host, so the per-tick
    graph launch (~4 us of launch gap at 4096 envs) is paid once per rollout instead.  Every engine must advance by a
    multiple of its number of output sets per replay, so that the captured set bindings stay valid."""

    def __init__(self, engines, tp_weights, ticks: int, raw: bool = True):
        n = len(engines)
        if ticks % n != 0 or any((ticks // n) % len(e.sets) != 0 for e in engines):
            raise _lib.HsError("RotatingRolloutGraph: ticks must be a multiple of len(engines) x output sets per engine")
        self.engines, self.ticks, self.per_engine = list(engines), ticks, ticks // n
        dev = engines[0].device
        for e in engines:
            if getattr(e, "graph_action", None) is None:
                e.graph_action = torch.zeros(e.E, e.A, 4, dtype=torch.float32, device=dev)
                e.graph_reset_pid = torch.zeros(e.E, dtype=torch.uint8, device=dev)
        start = [e.cur for e in engines]
        counts0 = [int(lib.hs_launch_count(e._h)) for e in engines]
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        self.graph = torch.cuda.CUDAGraph()
        self._keep = list(tp_weights)
        with torch.cuda.graph(self.graph, stream=side):
            st = torch.cuda.current_stream(dev).cuda_stream
            for t in range(ticks):
                e, w = engines[t % n], tp_weights[t % n]
                prev, i = e.cur, e.next_index()
                action = e.graph_action
                if getattr(e, "_actor", None) is not None:       # attached policy: actor (+ critic) on the previous observation
                    e._launch_policy(prev, i)
                    action = e.policy_out[i]["action"]
                    self.policy_kernels = 2 if e._critic is not None else 1
                e._bind(i, prev)
                if e.cfg.use_tp_net:
                    check(lib.hs_step_fused(e._h, action.data_ptr(), 1 if raw else 0, e.graph_reset_pid.data_ptr(),
                                            C.byref(w), None, st), "hs_step_fused (capture)")
                else:
                    check(lib.hs_step_pre(e._h, action.data_ptr(), 1 if raw else 0, e.graph_reset_pid.data_ptr(), st),
                          "hs_step_pre (capture)")
        # kernels per replay and engine; the capture-time calls were counted by the handle but never ran
        self.kernels = [int(lib.hs_launch_count(e._h)) - c for e, c in zip(engines, counts0)]
        for e, c, k in zip(engines, start, self.kernels):
            e._bind(c, c)
            e._uncounted = getattr(e, "_uncounted", 0) - k
        self.replays = 0

    def replay(self):
        self.graph.replay()
        self.replays += 1
        for e in self.engines:
            if e.host_max_progress is not None:
                e.host_max_progress += self.per_engine
        for e, k in zip(self.engines, self.kernels):
            # (cur is unchanged: every engine advanced by a multiple of its set count)
            e._uncounted = getattr(e, "_uncounted", 0) + k + getattr(self, "policy_kernels", 0) * self.per_engine


class PolicyRolloutGraph:
    """ONE CUDA graph = one whole rollout of a rollout-mode engine with an attached policy
    (``attach_policy(actor, critic, defer_critic=True)``): T x (actor on the latest observation -> tick + predictor into
    row t of the time-major storage), then the critic ONCE over the T stored observations.  The reference's collector
    evaluates the critic every step (``MAPPOPolicy.__call__`` -> ``value_op``, mappo.py:235-251); it has no state, so one
    launch over T x E x A rows gives the same values at a fraction of the cost of T small launches.
    After :meth:`replay`: ``storage.policy["state_value"][t]`` = V(observation step t started from) and
    ``next_state_value[t]`` = V(observation step t produced) - the two inputs of ``compute_gae``."""

    def __init__(self, engine, tp_weights, raw: bool = True):
        e = engine
        if e.storage is None or getattr(e, "_actor", None) is None or not getattr(e, "_defer_critic", False):
            raise _lib.HsError("PolicyRolloutGraph needs rollout mode and attach_policy(actor, critic, defer_critic=True)")
        T, dev = e.storage.T, e.device
        self.engine, self.T = e, T
        obs_keys = ("state_self", "state_others", "obs_cylinders")
        if any(not e.storage.data[k][:T].is_contiguous() for k in obs_keys if not (k == "state_others" and e.A == 1)):
            raise _lib.HsError("PolicyRolloutGraph: the rollout rows of the observation are padded (num_envs x row width not a multiple of 32 words)")
        self.next_state_value = torch.zeros(T, e.E, e.A, 1, dtype=torch.float32, device=dev)
        if getattr(e, "graph_reset_pid", None) is None:
            e.graph_reset_pid = torch.zeros(e.E, dtype=torch.uint8, device=dev)
        # steady state: the latest observation sits in row T-1 (a first rollout is run eagerly to get there)
        while e._slot != T - 1:
            e.policy_tick(tp_weights)
        self._critic_rows(T - 1, T)                              # carry for the first replay
        sv = e.storage.policy["state_value"]
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        self.graph = torch.cuda.CUDAGraph()
        self._keep = tp_weights
        n0 = int(lib.hs_launch_count(e._h))
        with torch.cuda.graph(self.graph, stream=side):
            st = torch.cuda.current_stream(dev).cuda_stream
            sv[0].copy_(self.next_state_value[T - 1])           # value of the observation step 0 starts from
            for t in range(T):
                prev, i = e.cur, e.next_index()
                e._launch_policy(prev, i)
                e._bind(i, prev)
                e._slot = i
                check(lib.hs_step_fused(e._h, e.policy_out[i]["action"].data_ptr(), 1 if raw else 0,
                                        e.graph_reset_pid.data_ptr(), C.byref(tp_weights), None, st), "hs_step_fused (capture)")
            self._critic_rows(0, T)
            sv[1:T].copy_(self.next_state_value[:T - 1])
        self.kernels = int(lib.hs_launch_count(e._h)) - n0
        e._uncounted = getattr(e, "_uncounted", 0) - self.kernels
        self.policy_kernels_per_rollout = T + 1
        self.replays = 0

    def _critic_rows(self, a: int, b: int):
        e, d = self.engine, self.engine.storage.data
        others = d["state_others"][a:b] if e.A > 1 else None
        e._critic.forward(d["state_self"][a:b], others, d["obs_cylinders"][a:b], out={"head": self.next_state_value[a:b]})

    def replay(self):
        e = self.engine
        self.graph.replay()
        self.replays += 1
        if e.host_max_progress is not None:
            e.host_max_progress += self.T
        e._uncounted = getattr(e, "_uncounted", 0) + self.kernels + self.policy_kernels_per_rollout


class HostIoLoop:
    """hs_step_host_io_many: K host-buffer ticks rotating over several engines (independent env batches of one GPU) from
    ONE C call - per tick pinned host action in, tick + predictor, observation / reward / done back in pinned host memory,
    `in_flight` batches between issue and wait.  The per-tick host work of the Python loop around
    ``HsEngine.step_host(sync=False)`` / ``wait_host()`` (ctypes call, graph lookup, set rotation, stream sync: ~25 us,
    and N ranks share the host's cores) stays inside the library.  ``on_obs(batch_index)`` is called when a batch's
    results are in host memory - where a host-side policy writes that batch's next action."""

    def __init__(self, engines, tp_weights, actions_host, streams=None):
        dev = engines[0].device
        n = len(engines)
        self.engines, self.weights = list(engines), tp_weights
        self.streams = streams or [torch.cuda.Stream(dev) for _ in range(2)]
        self.batches = (_lib.hs_host_batch * n)()
        self._keep = [actions_host]
        self.views = []
        for b, (e, act) in enumerate(zip(engines, actions_host)):
            if e.storage is not None:
                raise _lib.HsError("HostIoLoop needs slab output sets (rollout_steps=None)")
            E, A = e.E, e.A
            ns = len(e.sets)
            mirror = torch.empty(e.sets[0].policy_words, dtype=torch.float32).pin_memory()
            done = torch.empty(E, dtype=torch.uint8).pin_memory()
            staging = torch.empty(E, A, 4, dtype=torch.float32, device=dev)
            sets = (hs_buffers * ns)(*[e._make_bufs(i) for i in range(ns)])
            ios = (_lib.hs_host_io * ns)()
            per_set = []
            for i in range(ns):
                out, base = e.sets[i], e.sets[i].slab.data_ptr()
                v = {}
                for k in ("state_self", "state_others", "obs_cylinders", "reward"):
                    t = out.t[k]
                    off = (t.data_ptr() - base) // 4
                    v[k] = mirror[off:off + t.numel()].view(t.shape)
                    setattr(ios[i], k, v[k].data_ptr() if t.numel() else None)
                ios[i].done = done.data_ptr()
                ios[i].action = act.data_ptr()
                per_set.append(v)
            self.views.append((per_set, done))
            hb = self.batches[b]
            hb.h, hb.sets, hb.ios, hb.num_sets = e._h, sets, ios, ns
            hb.next_set = e.next_index()
            hb.staging_dev = staging.data_ptr()
            hb.stream = self.streams[b % len(self.streams)].cuda_stream
            self._keep += [mirror, done, staging, sets, ios]

    def run(self, num_ticks: int, in_flight: int = 2, on_obs=None):
        cb = _lib.HS_OBS_CALLBACK(lambda user, b: on_obs(int(b))) if on_obs is not None else None
        w = self.weights
        for b, e in enumerate(self.engines):
            self.batches[b].next_set = e.next_index()
        check(lib.hs_step_host_io_many(self.batches, len(self.engines), int(num_ticks), int(in_flight), 1,
                                       C.byref(w) if w is not None else None, cb, None), "hs_step_host_io_many")
        n = len(self.engines)
        for b, e in enumerate(self.engines):
            ticks_b = num_ticks // n + (1 if b < num_ticks % n else 0)
            if ticks_b:
                e.cur = (self.batches[b].next_set - 1) % len(e.sets)      # the library bound that set last
                if e.host_max_progress is not None:
                    e.host_max_progress += ticks_b
        return self
