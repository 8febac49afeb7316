"""The C++ host mirror (ipc-toolkit_b200/cpp/ipcb200.hpp) compiles with g++ against the C ABI and,
on a GPU, reproduces the reference's codim known-answer test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")
LIBDIR = os.path.join(ROOT, "ipc-toolkit_b200")


def _build():
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", SRC, "-o", EXE, "-L" + LIBDIR, "-lipcb200", "-Wl,-rpath," + LIBDIR,
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_known_answers():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout
