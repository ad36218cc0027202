"""omni_drones/utils/torchrl/__init__.py of the reference: collector + agent spec."""
from mupe_b200.compat import SyncDataCollector  # noqa: F401
from mupe_b200.envs.agent_spec import AgentSpec  # noqa: F401
