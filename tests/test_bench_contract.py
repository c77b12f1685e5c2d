"""CPU: the reference arm of bench.py (the unmodified reference from oracle/_ref on the host cores) prints the
JSON line the driver expects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "1", "--cpu_sites", "200"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = p.stdout.strip().splitlines()[-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "ind-sites/sec per EM iteration" and d["unit"] == "ind-sites/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.ref
def test_reference_arm_sizes_its_sample_from_a_time_budget():
    """Without --cpu_sites the sample is sized from one calibration iteration so that W + K iterations fit the
    budget (the driver chooses K and W); the floor is 2,000 sites."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--cpu_budget_s", "0.5"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert "100 individuals x 2000 sites" in d["cpu_baseline"]["sample"]
    assert "calibration iteration" in d["cpu_baseline"]["sample"]
    assert d["config"]["workload"].startswith("configs[1] per GPU: 100 individuals x 1000000 sites")
