"""Import alias.  The package directory is ``revs-admm_b200/`` (the project's name); a
hyphen is not a Python identifier, so ``import revs_admm_b200`` loads that directory."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "revs-admm_b200")
_spec = importlib.util.spec_from_file_location(
    "revs_admm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["revs_admm_b200"] = _mod
_spec.loader.exec_module(_mod)
