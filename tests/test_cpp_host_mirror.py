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


SHARD_SRC = os.path.join(ROOT, "tests", "cpp", "test_sharded_step.cpp")
SHARD_EXE = os.path.join(ROOT, "tests", "cpp", "test_sharded_step")


def _build_sharded():
    if not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("no NCCL headers")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I/usr/local/cuda/include", SHARD_SRC, "-o", SHARD_EXE, "-L" + LIBDIR, "-lipcb200",
           "-Wl,-rpath," + LIBDIR, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_sharded_step_compiles_and_links():
    """ipc-toolkit_b200/cpp/ipcb200_sharded.hpp: the multi-GPU contact step in C++ over libipcb200 + NCCL"""
    _build_sharded()
    assert os.path.exists(SHARD_EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 2])
def test_cpp_sharded_step_nccl(lanes):
    """one process, one host thread per GPU (ncclCommInitAll): the ranks reproduce the single-GPU step (energy, gradient,
    step size after the all-reduces; row blocks tiling the single-GPU Hessian)"""
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    _build_sharded()
    r = subprocess.run([SHARD_EXE, str(world), str(lanes - 1)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded step ok" in r.stdout
