"""omni_drones/learning of the reference (MAPPO / PPO learners) is the CONSUMER of this path and out of scope
(SURVEY.md section 2 row 13).  When the reference's own package is importable (its dependencies tensordict / torchrl
installed and a copy of the reference on the path) the scripts should import that one - do not put this shim ahead of
it.  What is here: an inference-only `MAPPOPolicy` on the fused actor / critic kernels (hs_policy_forward) with the
constructor and call signature scripts/train.py uses, so that the collection loop of the script runs end to end;
`train_op` does not learn and says so."""
import warnings

import torch

from mupe_b200.policy import FusedPolicy, MAPPOActorCritic, init_params


class MAPPOPolicy:
    def __init__(self, cfg, agent_spec, device="cuda", TP_net=None):
        self.cfg, self.agent_spec, self.TP = cfg, agent_spec, TP_net
        dev = torch.device(device if str(device) != "cuda" else "cuda:0")
        obs = agent_spec.observation_spec
        d_self = int(obs["state_self"].shape[-1])
        n_others = int(obs["state_others"].shape[-2]) if "state_others" in obs else 0
        n_cyl = int(obs["cylinders"].shape[-2]) if "cylinders" in obs else 0
        self.actor_params = init_params(d_self, n_others, n_cyl, 4, True, dev)
        self.critic_params = init_params(d_self, n_others, n_cyl, 1, False, dev)
        self._ac = MAPPOActorCritic(FusedPolicy(self.actor_params, n_others, n_cyl, dev).seed(int(getattr(cfg, "seed", 0) or 0)),
                                    FusedPolicy(self.critic_params, n_others, n_cyl, dev), agent_name=agent_spec.name)
        self._warned = False

    def __call__(self, tensordict, deterministic: bool = False):
        return self._ac(tensordict, deterministic=deterministic)

    def train_op(self, tensordict):
        if not self._warned:
            warnings.warn("omni_drones.learning shim: MAPPOPolicy.train_op does not learn (the learner is outside the B200 "
                          "env tier); install the reference's learning package to train")
            self._warned = True
        return {}

    def state_dict(self):
        return {"actor_params": self.actor_params, "critic": self.critic_params,
                "TP": self.TP.state_dict() if self.TP is not None else {}}

    def load_state_dict(self, sd):
        for k, v in sd.get("actor_params", {}).items():
            self.actor_params[k].copy_(v)
        for k, v in sd.get("critic", {}).items():
            self.critic_params[k].copy_(v)
        self._ac.refresh()


Policy = PPOPolicy = PPOAdaptivePolicy = PPORNNPolicy = MAPPOPolicy
