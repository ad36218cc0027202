from .agent_spec import AgentSpec
from .hideandseek import HideAndSeek, PIDRateController
from .isaac_env import IsaacEnv
from .tp_net import TP_net

__all__ = ["AgentSpec", "HideAndSeek", "PIDRateController", "IsaacEnv", "TP_net"]
