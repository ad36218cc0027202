"""omni_drones/envs/isaac_env.py of the reference -> the B200 base class and its REGISTRY."""
from mupe_b200.envs import HideAndSeek, HideAndSeek_envgen, Hover  # noqa: F401  (registers the task classes)
from mupe_b200.envs.agent_spec import AgentSpec  # noqa: F401
from mupe_b200.envs.isaac_env import IsaacEnv  # noqa: F401
