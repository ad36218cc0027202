"""B200-native HideAndSeek environment step (drop-in for the reference's IsaacEnv path).

Importing this package loads ``libhs_b200.so`` (the sm_100a kernels behind the C ABI of
``include/hs_b200.h``).  There is no CPU implementation: a missing library is an
ImportError, a missing GPU is an ``HsError`` at environment construction.
"""
from . import _lib
from ._lib import HsError
from .compat import (Compose, InitTracker, SyncDataCollector, TensorDict, TransformedEnv, step_mdp)
from .config import Cfg, build_hs_config, compose, load_drone_params
from .engine import HsEngine, RolloutStorage
from . import rollout
from .rollout import compute_gae
from .policy import FusedPolicy, MAPPOActorCritic
from . import parallel
from .envs import AgentSpec, GenBuffer, HideAndSeek, HideAndSeek_envgen, Hover, IsaacEnv, PIDRateController, TP_net

__all__ = ["parallel", "rollout", "compute_gae", "RolloutStorage", "FusedPolicy", "MAPPOActorCritic", "HsError", "HsEngine", "Cfg", "build_hs_config", "compose", "load_drone_params", "TensorDict",
           "TransformedEnv", "Compose", "InitTracker", "SyncDataCollector", "step_mdp", "AgentSpec",
           "HideAndSeek", "HideAndSeek_envgen", "GenBuffer", "Hover", "IsaacEnv", "PIDRateController", "TP_net"]
