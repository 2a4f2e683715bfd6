"""bench.py --impl reference: the CPU arm the driver runs beside the GPU arm.  No GPU involved: the reference's own translation
units (oracle/_ref, when built) or the oracle port run the per-frame path on the host threads and the ONE result line carries the
contract's keys with the same `config` object the GPU arm prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                           "--batch", "4", "--pool", "8"] + extra, capture_output=True, text=True, timeout=600, env=e)


@pytest.mark.parametrize("workload", ["kitti", "tum"])
def test_reference_arm_prints_one_contract_line(workload, monkeypatch):
    r = run(["--workload", workload])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # the same config object as the GPU arm (bench.workload_config): workload string, frames per step, pool
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "POOL", 8)          # other tests read bench.POOL (the 256-frame sweep)
    assert d["config"] == bench.workload_config(workload, 4)
    assert ("RGB-D constructor" in d["config"]["stages"]) == (workload == "tum")


def test_reference_arm_other_ranks_exit_without_work():
    r = run([], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
