from mupe_b200.envs import HideAndSeek, HideAndSeek_envgen, Hover, IsaacEnv  # noqa: F401  (importing registers them)
