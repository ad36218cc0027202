"""omni_drones/utils/torchrl/transforms.py of the reference.  `PIDRateController` is the one transform on the hot path
(its arithmetic is fused into the tick kernel); the others are the alternative action / observation transforms that
SURVEY.md section 2 row 10b marks out of scope - their names exist so that `scripts/train.py` imports unmodified, and
they say so when a config selects them."""
from mupe_b200.envs.hideandseek import PIDRateController  # noqa: F401


def _out_of_scope(name):
    class _T:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: not part of the HideAndSeek hot path (tasks use action_transform: PIDrate)")
    _T.__name__ = name
    return _T


LogOnEpisode = _out_of_scope("LogOnEpisode")
FromMultiDiscreteAction = _out_of_scope("FromMultiDiscreteAction")
FromDiscreteAction = _out_of_scope("FromDiscreteAction")
History = _out_of_scope("History")
VelController = _out_of_scope("VelController")
AttitudeController = _out_of_scope("AttitudeController")
RateController = _out_of_scope("RateController")


def ravel_composite(*a, **k):
    raise NotImplementedError("ravel_composite: flatten_obs / flatten_state are not used by the HideAndSeek task files")
