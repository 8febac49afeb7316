"""The C-ABI library loads and exports every symbol include/ipcb200.h declares (no compute calls: no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(prefix_macro="IPCB_FN"):
    src = open(os.path.join(ROOT, "include", "ipcb200.h")).read()
    host, dev = src.split("#ifndef IPCB_ORACLE", 1)
    names = lambda s: sorted(set(re.findall(r"IPCB_FN\((\w+)\)\(", s)))
    return names(host), names(dev)


def test_header_and_binding_tables_agree():
    import ipctk_b200

    abi = ipctk_b200._pkg._abi
    host, dev = _declared()
    assert host == sorted(abi.HOST_API), set(host) ^ set(abi.HOST_API)
    assert dev == sorted(abi.DEVICE_API), set(dev) ^ set(abi.DEVICE_API)


def test_product_library_exports_every_symbol():
    import ipctk_b200

    pkg = ipctk_b200._pkg
    assert os.path.exists(pkg.LIB_PATH), "run `python __graft_entry__.py build` first"
    lib = pkg._abi.Lib(pkg.LIB_PATH, "ipcb_", device_api=True)  # raises AttributeError on a missing symbol
    assert lib.backend() == "cuda-sm100a"
    host, dev = _declared()
    assert lib.names == sorted(host + dev)


def test_oracle_exports_host_half(oracle):
    assert oracle.lib.backend() == "oracle-cpu"
    assert oracle.lib.names == _declared()[0]


def test_product_has_no_cpu_fallback():
    """without a CUDA device the product must fail loudly instead of computing on the host"""
    import ipctk_b200

    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    import numpy as np

    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ipctk_b200.CollisionMesh(np.zeros((3, 3)))


def test_product_sources_never_touch_the_oracle():
    pkg_dir = os.path.join(ROOT, "ipc-toolkit_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "ipco_" not in text, f
