"""omni_drones/utils/wandb.py::init_wandb of the reference (logging is outside the tier): a real wandb run when wandb is
importable and not disabled by the config, else a local stand-in with the attributes the scripts touch."""
import datetime
import os


class _LocalRun:
    def __init__(self, name, directory):
        self.name, self.dir = name, directory
        self.history = []

    def log(self, info):
        self.history.append(info)

    def finish(self):
        pass


def init_wandb(cfg):
    w = cfg.get("wandb", None) if hasattr(cfg, "get") else None
    mode = (w.get("mode", "disabled") if w is not None and hasattr(w, "get") else "disabled") or "disabled"
    stamp = datetime.datetime.now().strftime("%m-%d_%H-%M")
    name = f"{cfg.task.name}-{cfg.algo.name}/{stamp}"
    if mode != "disabled":
        try:
            import wandb
            return wandb.init(project=w.get("project", "omnidrones"), name=name, mode=mode)
        except Exception:
            pass
    d = os.path.join("outputs", name)
    os.makedirs(d, exist_ok=True)
    return _LocalRun(name, d)
