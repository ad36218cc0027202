"""IsaacEnv-compatible base class with the simulator replaced by the B200 engine.

Mirrors the public surface of the reference's base env (omni_drones/envs/isaac_env.py:47-260):
constructor ``(cfg, headless)``, the class REGISTRY (by name and lower-cased name),
``agent_spec``, ``reset()/step()`` through the torchrl EnvBase protocol with the ``_reset``
mask, ``progress_buf``, ``enable_render/render`` (no-ops: there is no viewport), ``close``.
Isaac Sim / PhysX are not involved: state lives in the engine's arena.
"""
import abc
from typing import Callable, Dict, Type, Union

import torch

from ..compat import EnvBase, TensorDict
from .agent_spec import AgentSpec


class _AgentSpecView(dict):
    def __init__(self, env):
        super().__init__(env._agent_spec)
        self.env = env

    def __setitem__(self, k: str, v: AgentSpec) -> None:
        v._env = self.env
        self.env._agent_spec[k] = v
        dict.__setitem__(self, k, v)


class IsaacEnv(EnvBase):
    REGISTRY: Dict[str, Type["IsaacEnv"]] = {}

    def __init__(self, cfg, headless: bool = True):
        device = cfg.sim.device if cfg.sim is not None and cfg.sim.device else "cuda:0"
        super().__init__(device=device, batch_size=[cfg.env.num_envs], run_type_checks=False)
        self.cfg = cfg
        self.enable_render(not headless)
        self.num_envs = cfg.env.num_envs
        self.max_episode_length = cfg.env.max_episode_length
        self.substeps = cfg.sim.substeps if cfg.sim is not None else 1
        self.dt = float(cfg.sim.dt) if cfg.sim is not None and cfg.sim.dt else 0.01
        self._is_closed = False
        self._design_scene()
        self._set_specs()

    @classmethod
    def __init_subclass__(cls, **kwargs):
        if cls.__name__ in IsaacEnv.REGISTRY:
            raise ValueError(f"duplicate env class {cls.__name__}")
        super().__init_subclass__(**kwargs)
        if not cls.__name__.startswith("_"):
            IsaacEnv.REGISTRY[cls.__name__] = cls
            IsaacEnv.REGISTRY[cls.__name__.lower()] = cls

    @property
    def agent_spec(self):
        if not hasattr(self, "_agent_spec"):
            self._agent_spec = {}
        return _AgentSpecView(self)

    @agent_spec.setter
    def agent_spec(self, value):
        raise AttributeError("Do not set agent_spec directly; use self.agent_spec[name] = AgentSpec(...)")

    @abc.abstractmethod
    def _set_specs(self):
        raise NotImplementedError

    @abc.abstractmethod
    def _design_scene(self):
        raise NotImplementedError

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError(f"Cannot move IsaacEnv on {self.device} to {device} once it is initialized.")
        return self

    def enable_render(self, enable: Union[bool, Callable] = True):
        self._should_render = (lambda substep: enable) if isinstance(enable, bool) else enable

    def render(self, mode: str = "human"):
        """No viewport in this backend: "rgb_array" returns a black 4x4 frame so that the reference's evaluation loop
        (scripts/train.py:212-246 stacks the frames into a video) keeps its shape contract."""
        if mode == "rgb_array":
            import numpy as np
            return np.zeros((4, 4, 3), dtype=np.uint8)
        return None

    def close(self):
        self._is_closed = True

    def _set_seed(self, seed=-1):
        torch.manual_seed(seed)
