"""B200-native HideAndSeek environment step (drop-in for the reference's IsaacEnv path).

Importing this package loads ``libhs_b200.so`` (the sm_100a kernels behind the C ABI of
``include/hs_b200.h``).  There is no CPU implementation: a missing library is an
ImportError, a missing GPU is an ``HsError`` at environment construction.
"""
from . import _lib
from ._lib import HsError
from .compat import (Compose, InitTracker, SyncDataCollector, TensorDict, TransformedEnv, step_mdp)
from .config import Cfg, build_hs_config, compose, load_drone_params
from .engine import HostIoLoop, HsEngine, RolloutStorage
from . import rollout
from .rollout import compute_gae, gather_minibatch, make_dataset_naive
from .policy import FusedPolicy, MAPPOActorCritic
from . import parallel
from .envs import AgentSpec, GenBuffer, HideAndSeek, HideAndSeek_envgen, Hover, IsaacEnv, PIDRateController, TP_net



def shim_path() -> str:
    """Directory holding the ``omni_drones`` import shim (multi-uav-pursuit-evasion_b200/shim/README.md)."""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def install_shim(force_standins: bool = False) -> dict:
    """Makes the reference's entry points importable against this backend: appends the ``omni_drones`` shim to the END of
    ``sys.path`` and, when the real ``tensordict`` / ``torchrl`` packages are absent (they are from this image; SURVEY.md
    fact 3), registers the in-repo stand-ins under those module names.  Real packages always win unless
    ``force_standins``.  Returns what was bound, e.g. ``{"tensordict": "stand-in", "torchrl": "stand-in"}``."""
    import importlib.util
    import sys
    import types
    from . import compat
    if shim_path() not in sys.path:
        sys.path.append(shim_path())
    bound = {}
    for name in ("tensordict", "torchrl"):
        real = (not force_standins) and name not in sys.modules and importlib.util.find_spec(name) is not None
        if real or (name in sys.modules and not getattr(sys.modules[name], "__mupe_standin__", False) and not force_standins):
            bound[name] = "installed package"
            continue
        bound[name] = "stand-in"

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__mupe_standin__ = True
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    if bound["tensordict"] == "stand-in":
        mod("tensordict", TensorDict=compat.TensorDict, TensorDictBase=compat.TensorDictBase)
        mod("tensordict.tensordict", TensorDict=compat.TensorDict, TensorDictBase=compat.TensorDictBase)
    if bound["torchrl"] == "stand-in":
        specs = dict(CompositeSpec=compat.CompositeSpec, UnboundedContinuousTensorSpec=compat.UnboundedContinuousTensorSpec,
                     BoundedTensorSpec=compat.BoundedTensorSpec, DiscreteTensorSpec=compat.DiscreteTensorSpec,
                     TensorSpec=compat.TensorSpec)
        tr = dict(TransformedEnv=compat.TransformedEnv, InitTracker=compat.InitTracker, Compose=compat.Compose,
                  Transform=compat.Transform)
        mod("torchrl")
        mod("torchrl.data", **specs)
        mod("torchrl.envs", EnvBase=compat.EnvBase, step_mdp=compat.step_mdp, **tr)
        mod("torchrl.envs.transforms", **tr)
        mod("torchrl.envs.utils", step_mdp=compat.step_mdp)
        mod("torchrl.collectors", SyncDataCollector=compat.SyncDataCollector)
    return bound


__all__ = ["gather_minibatch", "make_dataset_naive", "shim_path", "install_shim", "HostIoLoop", "parallel", "rollout", "compute_gae", "RolloutStorage", "FusedPolicy", "MAPPOActorCritic", "HsError", "HsEngine", "Cfg", "build_hs_config", "compose", "load_drone_params", "TensorDict",
           "TransformedEnv", "Compose", "InitTracker", "SyncDataCollector", "step_mdp", "AgentSpec",
           "HideAndSeek", "HideAndSeek_envgen", "GenBuffer", "Hover", "IsaacEnv", "PIDRateController", "TP_net"]
