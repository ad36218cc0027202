"""Import shim: makes the package directory ``multi-uav-pursuit-evasion_b200/`` (whose name is
not a valid Python identifier) importable as ``mupe_b200``.

    import mupe_b200
    env = mupe_b200.HideAndSeek(cfg, headless=True)
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-uav-pursuit-evasion_b200")
_spec = importlib.util.spec_from_file_location(
    "mupe_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mupe_b200"] = _mod
_spec.loader.exec_module(_mod)
