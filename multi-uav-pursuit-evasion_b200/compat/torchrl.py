"""Minimal stand-in for the slice of torchrl 0.1.1 (pinned fork btx0424/rl@e39e701) that the
reference's env, transforms, collector and scripts/train.py use:
tensor specs, EnvBase.reset/step/rollout, Transform/Compose/TransformedEnv/InitTracker,
step_mdp and a synchronous collector.  Boundary glue only -- no hot-path arithmetic.
"""
import time
from typing import Callable, Dict, Iterator, List, Optional, Sequence

import torch

from .tensordict import TensorDict, _norm_key


# ------------------------------------------------------------------------------------------
# specs
# ------------------------------------------------------------------------------------------
class TensorSpec:
    def __init__(self, shape, device=None, dtype=torch.float32):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = torch.Size(shape)
        self.device = torch.device(device) if device is not None else None
        self.dtype = dtype

    def _like(self, shape):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.shape = torch.Size(shape)
        return out

    def expand(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        if len(shape) >= len(self.shape) and tuple(shape[len(shape) - len(self.shape):]) == tuple(self.shape):
            return self._like(shape)                       # full target shape given
        return self._like(tuple(shape) + tuple(self.shape))  # leading dims given

    def unsqueeze(self, dim):
        s = list(self.shape)
        s.insert(dim if dim >= 0 else len(s) + 1 + dim, 1)
        return self._like(s)

    def to(self, device):
        out = self._like(self.shape)
        out.device = torch.device(device)
        return out

    def zero(self, shape=()):
        return torch.zeros(*shape, *self.shape, dtype=self.dtype, device=self.device)

    def rand(self, shape=()):
        return torch.randn(*shape, *self.shape, device=self.device).to(self.dtype)

    def clone(self):
        return self._like(self.shape)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.stack:                       # torch.stack([spec] * n, 0)
            specs, dim = args[0], (args[1] if len(args) > 1 else kwargs.get("dim", 0))
            s = list(specs[0].shape)
            s.insert(dim, len(specs))
            return specs[0]._like(s)
        raise NotImplementedError(func)

    def __repr__(self):
        return f"{self.__class__.__name__}(shape={tuple(self.shape)}, device={self.device})"


class UnboundedContinuousTensorSpec(TensorSpec):
    pass


class BoundedTensorSpec(TensorSpec):
    def __init__(self, minimum, maximum, shape, device=None, dtype=torch.float32):
        super().__init__(shape, device, dtype)
        self.minimum, self.maximum = minimum, maximum

    def rand(self, shape=()):
        u = torch.rand(*shape, *self.shape, device=self.device)
        return (self.minimum + (self.maximum - self.minimum) * u).to(self.dtype)


class DiscreteTensorSpec(TensorSpec):
    def __init__(self, n, shape=(1,), device=None, dtype=torch.bool):
        super().__init__(shape, device, dtype)
        self.n = n

    def rand(self, shape=()):
        return torch.randint(0, self.n, (*shape, *self.shape), device=self.device).to(self.dtype)


class CompositeSpec(TensorSpec):
    def __init__(self, source: Dict = None, shape=(), device=None, **kw):
        super().__init__(shape, device)
        self._specs: Dict[str, TensorSpec] = {}
        for k, v in {**(source or {}), **kw}.items():
            self[k] = v

    def _like(self, shape):
        out = CompositeSpec({}, shape, self.device)
        nb_old, extra = len(self.shape), tuple(shape)
        for k, v in self._specs.items():
            out._specs[k] = v._like(extra + tuple(v.shape[nb_old:]))
        return out

    def expand(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        if len(shape) >= len(self.shape) and len(self.shape) and tuple(shape[len(shape) - len(self.shape):]) == tuple(self.shape):
            return self._like(shape)
        return self._like(tuple(shape) + tuple(self.shape))

    def to(self, device):
        out = CompositeSpec({}, self.shape, device)
        for k, v in self._specs.items():
            out._specs[k] = v.to(device)
        return out

    def clone(self):
        return self._like(self.shape)

    def __setitem__(self, key, value):
        key = _norm_key(key)
        if len(key) > 1:
            if key[0] not in self._specs:
                self._specs[key[0]] = CompositeSpec({}, self.shape, self.device)
            self._specs[key[0]][key[1:]] = value
            return
        if isinstance(value, dict):
            value = CompositeSpec(value, self.shape, self.device)
        self._specs[key[0]] = value

    def __getitem__(self, key):
        cur = self
        for k in _norm_key(key):
            cur = cur._specs[k]
        return cur

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def keys(self, include_nested=False, leaves_only=False) -> List:
        out = []
        for k, v in self._specs.items():
            if isinstance(v, CompositeSpec):
                if not leaves_only:
                    out.append(k)
                if include_nested:
                    out.extend((k,) + (s if isinstance(s, tuple) else (s,)) for s in v.keys(True, leaves_only))
            else:
                out.append(k)
        return out

    def items(self, include_nested=False, leaves_only=False):
        return [(k, self[k]) for k in self.keys(include_nested, leaves_only)]

    def _make(self, fn, shape=()):
        td = TensorDict({}, tuple(shape) + tuple(self.shape), self.device)
        for k, v in self._specs.items():
            td.set(k, v._make(fn, shape) if isinstance(v, CompositeSpec) else fn(v, shape))
        return td

    def zero(self, shape=()):
        return self._make(lambda s, sh: s.zero(sh), shape)

    def rand(self, shape=()):
        return self._make(lambda s, sh: s.rand(sh), shape)

    def __repr__(self):
        return "CompositeSpec(" + ", ".join(f"{k}: {v!r}" for k, v in self._specs.items()) + f", shape={tuple(self.shape)})"


# ------------------------------------------------------------------------------------------
# env base
# ------------------------------------------------------------------------------------------
def step_mdp(td: TensorDict, exclude_action: bool = False) -> TensorDict:
    """Root of the next tick = everything under "next" + the carried-over root entries."""
    nxt = td.get("next")
    out = td.exclude("next")              # a structural copy already (leaves shared, containers new)
    out.update(nxt)                       # containers that come from "next" are copied by update()
    return out


class EnvBase:
    def __init__(self, device="cpu", batch_size=(), run_type_checks=False):
        self.device = torch.device(device)
        self.batch_size = torch.Size(batch_size)
        self.input_spec = CompositeSpec({"_action_spec": CompositeSpec({}, self.batch_size)}, self.batch_size)
        self.output_spec = CompositeSpec({"_observation_spec": CompositeSpec({}, self.batch_size),
                                          "_reward_spec": CompositeSpec({}, self.batch_size)}, self.batch_size)
        self.done_spec = DiscreteTensorSpec(2, (*self.batch_size, 1), device=self.device)
        self.training = True

    # spec properties (the reference assigns them in _set_specs)
    observation_spec = property(lambda s: s.output_spec["_observation_spec"],
                                lambda s, v: s.output_spec.__setitem__("_observation_spec", v))
    action_spec = property(lambda s: s.input_spec["_action_spec"],
                           lambda s, v: s.input_spec.__setitem__("_action_spec", v))
    reward_spec = property(lambda s: s.output_spec["_reward_spec"],
                           lambda s, v: s.output_spec.__setitem__("_reward_spec", v))

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def set_seed(self, seed: Optional[int] = None):
        self._set_seed(seed)
        return seed

    def _set_seed(self, seed):
        if seed is not None:
            torch.manual_seed(seed)

    def close(self):
        pass

    def to(self, device):
        return self

    def fake_tensordict(self) -> TensorDict:
        td = self.observation_spec.zero()
        td.update(self.action_spec.zero())
        nxt = self.observation_spec.zero()
        nxt.update(self.reward_spec.zero())
        nxt.set("done", self.done_spec.zero())
        td.set("done", self.done_spec.zero())
        td.set("next", nxt)
        return td

    def reset(self, tensordict: Optional[TensorDict] = None, **kwargs) -> TensorDict:
        out = self._reset(tensordict, **kwargs)
        if "done" not in out:
            out.set("done", torch.zeros(*self.batch_size, 1, dtype=torch.bool, device=self.device))
        if tensordict is not None:
            tensordict.update(out)
            return tensordict
        return out

    def step(self, tensordict: TensorDict) -> TensorDict:
        out = self._step(tensordict)
        tensordict.update(out)
        return tensordict

    def rand_step(self, tensordict: Optional[TensorDict] = None):
        if tensordict is None:
            tensordict = TensorDict({}, self.batch_size, self.device)
        tensordict.update(self.action_spec.rand())
        return self.step(tensordict)

    def rollout(self, max_steps: int, policy: Optional[Callable] = None, callback: Optional[Callable] = None,
                auto_reset: bool = True, break_when_any_done: bool = True, return_contiguous: bool = True,
                tensordict: Optional[TensorDict] = None) -> TensorDict:
        td = self.reset() if auto_reset else tensordict
        frames = []
        for _ in range(max_steps):
            td = policy(td) if policy is not None else td.update(self.action_spec.rand())
            td = self.step(td)
            frames.append(td.clone())                 # deep copy: the env reuses its output buffers
            if callback is not None:
                callback(self, td)
            done = td.get(("next", "done"))
            if break_when_any_done and bool(done.any()):
                break
            td = step_mdp(td)
            if not break_when_any_done and bool(done.any()):
                td.set("_reset", done.clone())
                td = self.reset(td)
        return TensorDict.stack(frames, len(self.batch_size))


# ------------------------------------------------------------------------------------------
# transforms
# ------------------------------------------------------------------------------------------
class Transform:
    def __init__(self, in_keys=None, out_keys=None, in_keys_inv=None, out_keys_inv=None):
        self.in_keys, self.in_keys_inv = in_keys or [], in_keys_inv or []
        self.parent = None

    def _call(self, td):          # applied to the env's output ("next")
        return td

    def _inv_call(self, td):      # applied to the env's input before _step
        return td

    def reset(self, td):
        return td

    def transform_input_spec(self, spec):
        return spec

    def transform_observation_spec(self, spec):
        return spec

    def set_parent(self, env):
        self.parent = env


class Compose(Transform):
    def __init__(self, *transforms):
        super().__init__()
        self.transforms = list(transforms)

    def _call(self, td):
        for t in self.transforms:
            td = t._call(td)
        return td

    def _inv_call(self, td):
        for t in reversed(self.transforms):
            td = t._inv_call(td)
        return td

    def reset(self, td):
        for t in self.transforms:
            td = t.reset(td)
        return td

    def transform_input_spec(self, spec):
        for t in reversed(self.transforms):
            spec = t.transform_input_spec(spec)
        return spec

    def transform_observation_spec(self, spec):
        for t in self.transforms:
            spec = t.transform_observation_spec(spec)
        return spec

    def set_parent(self, env):
        self.parent = env
        for t in self.transforms:
            t.set_parent(env)


class InitTracker(Transform):
    """Adds "is_init": true on the first tick after a reset."""

    def _call(self, td):
        done = td.get("done")
        z = getattr(self, "_zeros", None)
        if z is None or z.shape != done.shape or z.device != done.device:
            z = self._zeros = torch.zeros_like(done)      # read-only constant: one allocation, no launch per step
        td.set("is_init", z)
        return td

    def reset(self, td):
        mask = td.get("_reset", None)
        done = td.get("done")
        init = torch.ones_like(done) if mask is None else mask.reshape(done.shape).clone()
        td.set("is_init", init)
        return td


class TransformedEnv(EnvBase):
    def __init__(self, env: EnvBase, transform: Optional[Transform] = None):
        self.base_env = env
        self.transform = transform or Compose()
        self.device, self.batch_size, self.training = env.device, env.batch_size, True
        self.transform.set_parent(self)
        self.input_spec = self.transform.transform_input_spec(env.input_spec.clone())
        self.output_spec = env.output_spec.clone()
        self.output_spec["_observation_spec"] = self.transform.transform_observation_spec(
            env.output_spec["_observation_spec"].clone())
        self.done_spec = env.done_spec

    def __getattr__(self, name):      # num_envs, agent_spec, max_episode_length, ...
        if name in ("base_env", "transform"):
            raise AttributeError(name)
        return getattr(self.base_env, name)

    def _set_seed(self, seed):
        self.base_env._set_seed(seed)

    def train(self, mode=True):
        self.base_env.train(mode)
        self.training = mode
        return self

    def close(self):
        self.base_env.close()

    def _reset(self, tensordict=None, **kwargs):
        out = self.base_env._reset(tensordict, **kwargs)
        if "done" not in out:
            out.set("done", torch.zeros(*self.batch_size, 1, dtype=torch.bool, device=self.device))
        if tensordict is not None and "_reset" in tensordict:
            out.set("_reset", tensordict.get("_reset"))
        out = self.transform.reset(out)
        out.pop("_reset", None)
        return out

    def _step(self, tensordict):
        tensordict = self.transform._inv_call(tensordict)
        out = self.base_env._step(tensordict)
        out.set("next", self.transform._call(out.get("next")))
        return out


# ------------------------------------------------------------------------------------------
# collector (what omni_drones/utils/torchrl/collector.py subclasses)
# ------------------------------------------------------------------------------------------
class SyncDataCollector:
    def __init__(self, env, policy=None, frames_per_batch=None, total_frames=-1, device=None,
                 return_same_td=True, reset_when_done=True, **unused):
        self.env, self.policy = env, policy
        self.n_envs = env.batch_size[0]
        self.frames_per_batch = frames_per_batch or self.n_envs
        self.total_frames = total_frames
        self.return_same_td = return_same_td
        self.reset_when_done = reset_when_done
        self.split_trajs, self.postproc, self._exclude_private_keys = False, None, True
        self._td = env.reset()
        self._fps = 0.0
        self._frames = 0

    def rollout(self) -> TensorDict:
        start = time.perf_counter()
        base = getattr(self.env, "base_env", self.env)
        eng = getattr(base, "engine", None)
        steps = self.frames_per_batch // self.n_envs
        if eng is not None and getattr(eng, "storage", None) is not None and eng.storage.T == steps \
                and hasattr(base, "rollout_next_td"):
            out = self._rollout_into_storage(base, eng, steps)
        else:
            frames = []
            for _ in range(steps):
                td = self._policy_step()
                frames.append(td.clone())
                self._carry(td)
            out = TensorDict.stack(frames, 1)
        self._fps = out.numel() / (time.perf_counter() - start)
        return out

    def _policy_step(self) -> TensorDict:
        td = self._td
        td = self.policy(td) if self.policy is not None else td.update(self.env.action_spec.rand())
        return self.env.step(td)

    def _any_done(self, done) -> bool:
        """`done.any()` without a device sync whenever the env can rule it out on the host (progress counters)."""
        base = getattr(self.env, "base_env", self.env)
        eng = getattr(base, "engine", None)
        if eng is not None and hasattr(eng, "maybe_done"):
            if eng.host_max_progress is None:
                eng.refresh_host_progress()               # one sync after a partial reset / injected state
            if not eng.maybe_done():
                return False
        return bool(done.any())

    def _carry(self, td: TensorDict):
        done = td.get(("next", "done"))
        td = step_mdp(td)
        if self.reset_when_done and self._any_done(done):
            td.set("_reset", done.clone())
            td = self.env.reset(td)
            td.pop("_reset", None)
        self._td = td

    def _rollout_into_storage(self, base, eng, steps: int) -> TensorDict:
        """Rollout mode: the tick kernels write ``next`` straight into the
        engine's time-major ``[T, E, ...]`` RolloutStorage, so the per-step clone and the final stack of
        the generic path disappear; the step's input side (observation the policy saw, its outputs)
        is copied once into preallocated ``[T, E, ...]`` tensors.  The result is the same ``[E, T]``
        tensordict, as views."""
        E, dev = self.n_envs, base.device
        if eng.rollout_slot not in (-1, steps - 1):
            raise RuntimeError("rollout storage is mid-rollout: the env was stepped outside the collector")
        pre = getattr(self, "_pre", None)
        for t in range(steps):
            td = self._policy_step()
            nxt = td.get("next")
            if pre is None:
                # first step ever: allocate the input-side storage from what the policy produced
                pre = self._pre = {}
                tup = lambda k: k if isinstance(k, tuple) else (k,)
                for k in td.keys(True, True):
                    if tup(k)[0] == "next":
                        continue
                    v = td.get(k)
                    pre[k] = torch.empty((steps,) + tuple(v.shape), dtype=v.dtype, device=v.device)
                self._stats_T = torch.empty((steps,) + tuple(eng.stats.shape), device=dev)     # [T, 24, E]: one copy per step
                self._prev_action_T = torch.empty((steps,) + tuple(eng.prev_action.shape), device=dev)
                # entries of ``next`` that transforms add on top of the engine's outputs (is_init, ...)
                engine_keys = {tup(k) for k in base.rollout_next_td(self._stats_T, self._prev_action_T).keys(True, True)}
                for k in nxt.keys(True, True):
                    if tup(k) not in engine_keys:
                        v = nxt.get(k)
                        pre[("next",) + tup(k)] = torch.empty((steps,) + tuple(v.shape), dtype=v.dtype, device=v.device)
            # one multi-tensor copy for the whole input side, one for the 24 stats rows (the engine keeps them as [24, E])
            dsts = [buf[t] for buf in pre.values()] + [self._stats_T[t], self._prev_action_T[t]]
            srcs = [td.get(k) for k in pre.keys()] + [eng.stats, eng.prev_action]
            torch._foreach_copy_(dsts, srcs)
            self._carry(td)
        out = TensorDict({}, [E, steps], dev)
        out.set("next", base.rollout_next_td(self._stats_T, self._prev_action_T))
        for k, buf in pre.items():
            out.set(k, buf.transpose(0, 1))
        return out

    def iterator(self) -> Iterator[TensorDict]:
        while True:
            out = self.rollout()
            self._frames += out.numel()
            yield out
            if 0 < self.total_frames <= self._frames:
                break

    def __iter__(self):
        return self.iterator()

    def shutdown(self):
        self.env.close()
