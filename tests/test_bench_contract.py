"""bench.py contract checks that run without a GPU: the reference arm prints exactly ONE JSON line on
stdout with the keys the driver reads, and the B200 arm refuses to run without a device."""
from __future__ import annotations

import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "plf_sites_per_s" and d["unit"] == "sites/s" and d["dtype"] == "f32"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert set(d["config"]) >= {"workload"} and "model" not in d["config"]
    assert d["config"]["sites_per_step"] == d["config"]["total_sites"] == 1 << 20      # same work per step as our arm
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_b200_arm_prints_one_complete_json_line():
    """A small run of our arm on the GPU: one JSON line with every key the driver and the judge read."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--sites", str(2 << 20), "--steps", "5",
                        "--warmup", "3", "--e2e-steps", "2", "--side-small"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["scaling"] in ("weak", "strong")
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.3 < rf["frac"] < 1.3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                       # host round trip is PCIe bound, never the device rate
    assert d["gpu_launches"] == 5                        # one fused kernel per step, nothing else of ours
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["scaler_increment"] == (2 << 20) // 4
    # host-link yardstick beside the end-to-end number
    er = e["roofline"]
    assert er["bound"] == "pcie" and er["peak"] > 5 and 0.3 < er["frac"] < 1.2
    # every other BASELINE config and section-8f row rides in the same line, each with its own roofline
    side = d["side"]
    assert {"cfg2", "cfg4", "cfg5", "protein", "evaluate"} <= set(side)
    for row in (side["cfg2"]["cold"], side["cfg2"]["warm"], side["cfg4"]["cfg4a_write"], side["cfg5"]["dense_tips"],
                side["cfg5"]["code_tips"], side["protein"]["strict"], side["protein"]["fma"], side["evaluate"]):
        assert row["value"] > 0 and row["roofline"]["frac"] > 0 and row["roofline"]["algorithmic_bytes"] > 0
    assert side["cfg4"]["cfg4b_discard"]["value"] > 0 and side["evaluate"]["bitwise_reproducible"] is True
