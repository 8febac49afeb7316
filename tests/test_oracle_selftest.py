"""oracle/selftest: closed-form derivatives vs second-order autodiff, eigen-solver, PSD known answers."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_selftest_binary_passes(oracle):  # the fixture builds oracle/ if needed
    r = subprocess.run([os.path.join(ROOT, "oracle", "selftest")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "all checks passed" in r.stdout
