"""Import shim: the names the reference's scripts import from ``omni_drones`` (omni_drones/__init__.py:27-61), bound to
the B200 backend.  No Kit / Isaac Sim application is started."""
import importlib.util
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))     # multi-uav-pursuit-evasion_b200/
if "mupe_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("mupe_b200", os.path.join(_PKG, "__init__.py"),
                                                   submodule_search_locations=[_PKG])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["mupe_b200"] = _mod
    _spec.loader.exec_module(_mod)
import mupe_b200  # noqa: E402

CONFIG_PATH = os.path.join(os.path.dirname(_PKG), "cfg")          # scripts pass it to hydra.main(config_path=...)


class _NoSimulationApp:
    """What ``init_simulation_app`` returns in the reference is the Kit application; here nothing has to run."""

    def close(self):
        pass

    def update(self):
        pass


def init_simulation_app(cfg=None):
    return _NoSimulationApp()
