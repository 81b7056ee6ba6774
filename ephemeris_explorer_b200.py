"""Import shim: the package directory is named `ephemeris-explorer_b200` (not a valid Python identifier).
`import ephemeris_explorer_b200` loads that directory as a regular package under this name."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "ephemeris-explorer_b200"
_spec = importlib.util.spec_from_file_location(
    __name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
