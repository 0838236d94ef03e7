"""bench.py contract on the CPU: the reference arm (`--impl reference`: the oracle's simplex over a process pool)
prints one JSON line with the keys the driver reads, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--case", "case9",
                          "--steps", "2", "--warmup", "1", "--ref-sample", "2"], capture_output=True, text=True,
                         timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_json_line():
    line = _run().splitlines()[-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "scenarios/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["optimal"] == 4 and "workload" in d["config"] and d["scaling"] == "strong" and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
