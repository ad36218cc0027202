from .agent_spec import AgentSpec
from .hideandseek import HideAndSeek, PIDRateController
from .hideandseek_envgen import GenBuffer, HideAndSeek_envgen, farthest_point_sampling
from .hover import Hover
from .isaac_env import IsaacEnv
from .tp_net import TP_net

__all__ = ["AgentSpec", "HideAndSeek", "HideAndSeek_envgen", "GenBuffer", "farthest_point_sampling", "Hover", "PIDRateController", "IsaacEnv", "TP_net"]
