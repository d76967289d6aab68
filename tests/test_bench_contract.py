"""bench.py's JSON contract, on the arm that needs no GPU (--impl reference: the reference's CPU decoder)."""
import json
import os
import subprocess
import sys

import workloads as W


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(W.ROOT, "bench.py"), "--impl", "reference", "--workload", "surface_d5",
                        "--shots", "4000", "--steps", "2", "--warmup", "1", "--ref-seconds", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded shots/sec" and d["unit"] == "shots/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "surface code d=5" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(W.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
