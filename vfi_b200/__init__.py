"""Import shim: the package lives in ``video-frame-interpolation_b200/`` (a name Python cannot import), so this
module loads that directory as the package ``vfi_b200``."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "video-frame-interpolation_b200"
_spec = _u.spec_from_file_location(__name__, _real / "__init__.py", submodule_search_locations=[str(_real)])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
