"""Import shim: the package directory is named after the reference repository
(``psychoacoustic-adverserial-attacks_b200``), which is not a valid Python identifier, so
``import paa_b200`` loads that directory as the package ``paa_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "psychoacoustic-adverserial-attacks_b200")
_spec = importlib.util.spec_from_file_location(
    "paa_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["paa_b200"] = _mod
_spec.loader.exec_module(_mod)
