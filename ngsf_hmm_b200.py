"""Import shim: the package directory is named ``ngsf-hmm_b200`` (a hyphen is
not a valid Python identifier), so ``import ngsf_hmm_b200`` loads it from there."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "ngsf-hmm_b200")
_spec = importlib.util.spec_from_file_location(
    "ngsf_hmm_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ngsf_hmm_b200"] = _mod
_spec.loader.exec_module(_mod)
