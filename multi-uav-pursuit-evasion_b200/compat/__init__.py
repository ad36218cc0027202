"""Container / env-protocol layer the environment is written against.

The reference pins forks of tensordict and torchrl (0.1.x API) that are absent from this
image, so the in-repo stand-ins are used.  They implement the subset of
that API which the env, the transforms, the collector and scripts/train.py exercise.
"""
from .tensordict import TensorDict, TensorDictBase
from .torchrl import (BoundedTensorSpec, CompositeSpec, Compose, DiscreteTensorSpec, EnvBase, InitTracker,
                      SyncDataCollector, TensorSpec, Transform, TransformedEnv, UnboundedContinuousTensorSpec,
                      step_mdp)

__all__ = ["TensorDict", "TensorDictBase", "BoundedTensorSpec", "CompositeSpec", "Compose", "DiscreteTensorSpec",
           "EnvBase", "InitTracker", "SyncDataCollector", "TensorSpec", "Transform", "TransformedEnv",
           "UnboundedContinuousTensorSpec", "step_mdp"]
