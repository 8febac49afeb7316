"""ipc-toolkit_b200 — B200-native per-step contact pipeline behind the IPC Toolkit API.

The product is ``libipcb200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/ipcb200.h``); this package is the thin Python host mirror of the
reference's ``ipctk`` interface for that path.  There is no CPU fallback: using
any class without the built CUDA library or without a CUDA device raises.

Because the directory name carries a hyphen, import it through the
``ipctk_b200`` shim at the repository root (``import ipctk_b200 as ipctk``).
"""
import os as _os

from . import _abi, scenes  # noqa: F401
from .api import make_api as _make_api, edges_from_faces, PSDProjectionMethod, TightInclusionCCD, AdditiveCCD  # noqa: F401

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "libipcb200.so")
_ns = None


def library():
    """the loaded product library; raises loudly if it has not been built"""
    global _ns
    if _ns is None:
        if not _os.path.exists(LIB_PATH):
            raise RuntimeError(
                "ipc-toolkit_b200: %s is missing — run `python __graft_entry__.py build` "
                "(there is no CPU fallback)" % LIB_PATH)
        _ns = _make_api(_abi.Lib(LIB_PATH, "ipcb_", device_api=True))
    return _ns


def __getattr__(name):  # CollisionMesh, NormalCollisions, BarrierPotential, ... resolve lazily
    if name.startswith("__"):
        raise AttributeError(name)
    ns = library()
    try:
        return getattr(ns, name)
    except AttributeError:
        raise AttributeError("module 'ipc-toolkit_b200' has no attribute %r" % name) from None
