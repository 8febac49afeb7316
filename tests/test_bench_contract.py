"""bench.py's reference arm (`--impl reference`: the CPU restatement of the reference path, no GPU needed) prints exactly
one JSON line with the keys the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line(oracle):
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["metric"].startswith("ms per contact step")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and sum(d["counts"]["collisions"]) > 0
    # the reference arm times the SAME workload as the GPU arm (no crop, no scaling) and reports the steps it really ran
    for key in ("triangles", "vertices", "edges", "dhat", "psd", "ccd", "l2", "parallelism"):
        assert key in d["config"], key
    assert d["config"]["triangles"] == 10176 and "full workload" in d["cpu_baseline"]["sample"]
    assert d["steps"] == 1 and d["steps_requested"] == 1


def test_reference_arm_other_ranks_stay_silent(oracle):
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
