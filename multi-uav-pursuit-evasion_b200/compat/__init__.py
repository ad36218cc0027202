"""Container / env-protocol layer the environment is written against.

The reference pins forks of tensordict and torchrl (0.1.x API: btx0424/tensordict@6d8119c, btx0424/rl@e39e701) that are
absent from this image.  Binding rule:

* when BOTH real packages are importable and speak that 0.1.x API (`tensordict.__version__` / `torchrl.__version__` start
  with "0.1"), their classes are used - the environment then IS a torchrl `EnvBase` and plugs into an unmodified
  reference install (set ``MUPE_COMPAT=standin`` to force the stand-ins);
* otherwise (absent, or a modern incompatible release) the in-repo stand-ins below are used.  They implement the subset
  of that API which the env, the transforms, the collector and scripts/train.py exercise; ``mupe_b200.install_shim()``
  can register them under the names ``tensordict`` / ``torchrl`` so that the reference's scripts import unmodified.

``compat.BACKEND`` says which one is bound.
"""
import os


def _real_versions():
    out = {}
    for name in ("tensordict", "torchrl"):
        try:
            mod = __import__(name)
            if getattr(mod, "__mupe_standin__", False):
                return None
            out[name] = str(getattr(mod, "__version__", ""))
        except Exception:
            return None
    return out


_versions = None if os.environ.get("MUPE_COMPAT", "") == "standin" else _real_versions()
if _versions is not None and all(v.startswith("0.1") for v in _versions.values()):
    from tensordict import TensorDict, TensorDictBase                                             # noqa: F401
    from torchrl.collectors import SyncDataCollector                                               # noqa: F401
    from torchrl.data import (BoundedTensorSpec, CompositeSpec, DiscreteTensorSpec, TensorSpec,  # noqa: F401
                              UnboundedContinuousTensorSpec)
    from torchrl.envs import EnvBase                                                               # noqa: F401
    from torchrl.envs.transforms import Compose, InitTracker, Transform, TransformedEnv           # noqa: F401
    from torchrl.envs.utils import step_mdp                                                        # noqa: F401
    BACKEND = {"kind": "installed packages", **_versions}
else:
    from .tensordict import TensorDict, TensorDictBase                                             # noqa: F401
    from .torchrl import (BoundedTensorSpec, CompositeSpec, Compose, DiscreteTensorSpec, EnvBase, InitTracker,   # noqa: F401
                          SyncDataCollector, TensorSpec, Transform, TransformedEnv, UnboundedContinuousTensorSpec,
                          step_mdp)
    BACKEND = {"kind": "in-repo stand-ins", "found": _versions}

__all__ = ["TensorDict", "TensorDictBase", "BoundedTensorSpec", "CompositeSpec", "Compose", "DiscreteTensorSpec",
           "EnvBase", "InitTracker", "SyncDataCollector", "TensorSpec", "Transform", "TransformedEnv",
           "UnboundedContinuousTensorSpec", "step_mdp", "BACKEND"]
