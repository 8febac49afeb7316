"""Import shim: the package directory is named ``ipc-toolkit_b200`` (hyphen), so
``import ipctk_b200 as ipctk`` loads it by path and re-exports it."""
import importlib.util as _u
import os as _os
import sys as _sys

_name = "ipc_toolkit_b200"
if _name not in _sys.modules:
    _dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ipc-toolkit_b200")
    _spec = _u.spec_from_file_location(_name, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_pkg = _sys.modules[_name]


def __getattr__(name):
    return getattr(_pkg, name)
