"""bench.py's one-line JSON contract, checked on the committed bench lines (profiles/) and on bench.py's own
argument handling — no GPU needed."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

OURS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"]


def _lines(pattern):
    out = []
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        txt = open(f).read().strip()
        if txt:
            out.append((os.path.basename(f), json.loads(txt.splitlines()[-1])))
    return out


def test_latest_single_gpu_line_has_every_contract_key():
    lines = [(n, d) for n, d in _lines("r01_bench_n1_*.json") if "no_cpu_baseline" not in n]
    assert lines, "no committed bench line"
    name, d = lines[-1]
    for k in OURS:
        assert k in d, (name, k)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert d["dtype"] == "u8" and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                       # the host leg is PCIe-bound, never faster than the HBM number
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    lines = _lines("r01_bench_ref_*.json")
    assert lines
    name, d = lines[-1]
    assert d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    ours = [x for _, x in _lines("r01_bench_n1_*.json")][-1]
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"]


def test_multi_gpu_lines_scale_weakly():
    per_gpu = {}
    for name, d in _lines("r01_bench_n[0-9]*.json"):
        per_gpu.setdefault(d["n_gpus"], []).append(d["value"] / d["n_gpus"])
    assert 1 in per_gpu and max(per_gpu) >= 2
    base = max(per_gpu[1])
    for n, vals in per_gpu.items():
        assert max(vals) > 0.9 * base, (n, vals)         # device-resident copy: no cross-GPU dependency


def test_bench_help_and_bad_flag():
    ok = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "--impl" in ok.stdout + ok.stderr
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "nope"], capture_output=True, text=True,
                         timeout=120)
    assert bad.returncode != 0
