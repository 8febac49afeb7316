"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

Loads the CPU restatement (oracle/liboracle.so, prefix ipco_) behind the same
Python API classes as the product.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "ipc-toolkit_b200")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force=False):
    """compile liboracle.so / liboracle_fast.so / selftest with oracle/Makefile"""
    if force or not all(os.path.exists(os.path.join(HERE, f)) for f in ("liboracle.so", "liboracle_fast.so", "selftest")):
        subprocess.run(["make", "-C", HERE, "-j4"], check=True, stdout=subprocess.DEVNULL)


_abi = _load("_ipcb200_abi_for_oracle", os.path.join(PKG, "_abi.py"))
sys.modules.setdefault("_abi", _abi)
_api = _load("_ipcb200_api_for_oracle", os.path.join(PKG, "api.py"))


def load(fast=False):
    """returns the API namespace bound to the oracle (fast=True: the -mavx2 -mfma build)"""
    build()
    lib = _abi.Lib(os.path.join(HERE, "liboracle_fast.so" if fast else "liboracle.so"), "ipco_", device_api=False)
    ns = _api.make_api(lib)
    cd = lib.cdll
    cd.ipco_set_broad_method.argtypes = [C.c_void_p, C.c_int]
    cd.ipco_num_threads.restype = C.c_int
    cd.ipco_set_num_threads.argtypes = [C.c_int]
    cd.ipco_unit_distance_type.argtypes = [C.c_int, C.c_void_p]
    cd.ipco_unit_distance.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    cd.ipco_unit_mollifier.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    cd.ipco_unit_mollifier_threshold.argtypes = [C.c_void_p]
    cd.ipco_unit_mollifier_threshold.restype = C.c_double
    cd.ipco_unit_barrier.argtypes = [C.c_double, C.c_double, C.c_void_p]
    cd.ipco_unit_project_to_psd.argtypes = [C.c_int, C.c_void_p, C.c_int]
    cd.ipco_unit_morton_3D.argtypes = [C.c_double] * 3
    cd.ipco_unit_morton_3D.restype = C.c_uint64
    ns.cdll = cd
    ns.set_broad_method = lambda mesh, m: cd.ipco_set_broad_method(mesh._ctx, m)
    ns.num_threads = cd.ipco_num_threads
    ns.set_num_threads = cd.ipco_set_num_threads
    return ns
