"""Import shim: the package directory is `multirate.jl_b200/` (a dot in the name, as the task fixes it),
which Python cannot import by name.  `import multirate_b200 as mr` loads it as "multirate_jl_b200" and
re-exports its public names."""
import importlib.util
import os
import sys

_NAME = "multirate_jl_b200"


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkgdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multirate.jl_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkgdir, "__init__.py"),
                                                  submodule_search_locations=[pkgdir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load()
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
_ffi = _pkg._ffi
__all__ = list(_pkg.__all__)
