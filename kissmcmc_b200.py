"""Import shim: registers the package directory `kissmcmc.jl_b200/` (whose name is not a valid
Python identifier) as the importable module `kissmcmc_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg = Path(__file__).resolve().parent / "kissmcmc.jl_b200"
_spec = importlib.util.spec_from_file_location(__name__, _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
