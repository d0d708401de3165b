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


def test_round2_line_tells_the_whole_truth():
    """Round 2: the driver-run line carries all five BASELINE configs, each with its parity flag and a roofline object
    where a kernel dominates, a like-for-like file-tier e2e, and the tier its headline e2e was measured on."""
    lines = [(n, d) for n, d in _lines("r02_bench_n1_*.json")]
    assert lines, "no committed round-2 bench line"
    name, d = lines[-1]
    for k in OURS + ["e2e_file_tier", "extras"]:
        assert k in d, (name, k)
    assert "host_arena" in d["e2e"]["tier"] and d["e2e"]["pcie_probe"]["d2h_gbs"] > 0
    assert d["e2e"]["frac_of_pcie_probe"] <= 1.02                     # the probe is NUMA-local: a real ceiling
    assert len(d["e2e"]["store_gbs_per_gpu"]) == d["n_gpus"]
    ft = d["e2e_file_tier"]
    assert ft["bit_exact"] and ft["blocks_per_step"] >= 2048 and ft["value"] > 0 and "file" in ft["tier"]
    assert d["config"]["blocks_per_step"] == 10000
    assert d["cpu_baseline"]["blocks_per_step"] == ft["blocks_per_step"]      # like for like
    ex = d["extras"]
    for cfg in ("config1", "config3", "config5", "ingest", "manager_lookup"):
        assert cfg in ex, cfg
    assert ex["config1"]["bit_exact_vs_known_answer"] and ex["config1"]["us_per_call"] > 0
    c5 = ex["config5"]
    assert c5["bit_exact_vs_oracle"] and c5["index_keys"] >= 10_000_000 and c5["fused_tokens_to_scores_ms"] > 0
    assert c5["index_build"]["total_s"] < 2.0 and c5["index_build"]["flushes_parallel"] > 0     # built ON the device
    assert c5["hash_kernel"]["frac_of_floor"] > 0 and c5["cpu_baseline"]["kind"] == "port"
    c3 = ex["config3"]
    r = c3["roofline"]
    assert c3["bit_exact"] and r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert c3["e2e"]["bit_exact"] and c3["e2e"]["value"] > 0
    assert ex["ingest"]["bit_exact_vs_oracle"] and ex["manager_lookup"]["same_answer_as_the_loop"]
    # later in round 2: the targets the review set (one prompt <= 50 us, 1024 prompts within reach of 0.12 ms), small batches
    # and the index at capacity, each parity-asserted inside the bench before it reports a number
    assert ex["config1"]["us_per_call"] <= 50 and c5["fused_tokens_to_scores_ms"] <= 0.13
    sb = c5["small_batches"]
    assert set(sb) == {"4", "16", "64"} and all(v["bit_exact_vs_oracle"] and v["us_per_call"] > 0 for v in sb.values())
    cap = c5["index_at_capacity"]
    assert cap["lru_evictions"] == cap["new_keys"] == 200_000 and cap["plan_fallbacks"] == 0 and cap["seconds"] < 0.1
    mixed = cap["mixed_batch"]
    assert mixed["lru_evictions"] == 121_000 and mixed["flushes_planned"] == 1 and mixed["plan_fallbacks"] == 0
    burst = c5["hash_kernel"]["in_a_long_burst"]
    assert burst["us"] and burst["sm_mhz"] >= 0.9 * burst["sm_max_mhz"]


def test_round2_reference_arm_reports_what_it_moved():
    lines = _lines("r02_bench_ref_n*.json")
    if not lines:
        pytest.skip("no committed round-2 reference line yet")
    for name, d in lines:
        assert d["impl"] == "reference" and d["engines"] == d["n_gpus"]
        assert d["config"]["blocks_per_step"] <= d["config"]["workload_blocks"]
        assert str(d["config"]["blocks_per_step"]) in d["cpu_baseline"]["sample"]


def test_round2_multi_gpu_lines_verify_migration():
    lines = [(n, d) for n, d in _lines("r02_bench_n[2-8]*.json")]
    if not lines:
        pytest.skip("no committed round-2 multi-GPU line yet")
    for name, d in lines:
        mig = d["migration"]
        assert all(mig["ring"]["bit_exact"]) and len(mig["ring"]["bit_exact"]) == d["n_gpus"], name
        assert all(mig["ring_nccl_staged"]["bit_exact"]), name
        if d["n_gpus"] > 2:
            assert all(mig["fanout_rank0"]["bit_exact"]) and len(mig["fanout_rank0"]["bit_exact"]) == d["n_gpus"] - 1
        assert "traffic" in mig["ring"]["roofline"]
        assert "migration_70b" in d["extras"] and all(d["extras"]["migration_70b"]["ring"]["bit_exact"])
        assert d["e2e_file_tier"]["engines"] == d["n_gpus"]


def test_file_tier_blocks_are_bounded_and_honest():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.file_tier_blocks(5, 3) == 10000 // 16 * 16 or bench.file_tier_blocks(5, 3) >= 2048
    n = bench.file_tier_blocks(20, 5)
    assert 2048 <= n <= 10000 and n % 16 == 0
    assert bench.workload_config(1, n)["blocks_per_step"] == n and bench.workload_config(1, n)["workload_blocks"] == 10000


def test_bench_help_and_bad_flag():
    ok = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "--impl" in ok.stdout + ok.stderr
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "nope"], capture_output=True, text=True,
                         timeout=120)
    assert bad.returncode != 0
