"""AgentSpec: the view scripts/train.py and MAPPOPolicy read specs through
(reference: omni_drones/utils/torchrl/env.py:8-55)."""
from dataclasses import dataclass
from typing import Any, Optional


@dataclass
class AgentSpec:
    name: str
    n: int
    observation_key: Optional[Any] = "observation"
    action_key: Optional[Any] = None
    state_key: Optional[Any] = None
    reward_key: Optional[Any] = None
    done_key: Optional[Any] = None
    _env: Optional[Any] = None

    @property
    def observation_spec(self):
        return self._env.observation_spec[self.observation_key]

    @property
    def action_spec(self):
        if self.action_key is None:
            return self._env.action_spec
        try:
            return self._env.input_spec["_action_spec"][self.action_key]
        except KeyError:
            return self._env.action_spec[self.action_key]

    @property
    def state_spec(self):
        if self.state_key is None:
            raise ValueError("no state key")
        return self._env.observation_spec[self.state_key]

    @property
    def reward_spec(self):
        if self.reward_key is None:
            return self._env.reward_spec
        try:
            return self._env.output_spec["_reward_spec"][self.reward_key]
        except KeyError:
            return self._env.reward_spec[self.reward_key]

    @property
    def done_spec(self):
        return self._env.done_spec
