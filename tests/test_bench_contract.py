"""The bench.py JSON contract, checked on the lines committed under profiles/ (they were printed by bench.py on the
B200 box): every key the driver reads is present and well-formed.  CPU only."""
from __future__ import annotations

import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        pytest.skip(name + " not committed")
    return json.loads(open(p).read().strip().splitlines()[-1])


def _common(d):
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert isinstance(d["metric"], str) and d["unit"] == "events/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] >= 1 and d["steps"] >= 1
    assert d["scaling"] in ("weak", "strong") and d["data"] == "synthetic" and d["dtype"] == "u8"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric on this hardware
    assert isinstance(baseline, dict)
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and "h2d_bytes_per_step" in e and "d2h_bytes_per_step" in e


def test_b200_arm_line():
    d = _line("r01_final_bench.json")
    _common(d)
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("port", "reference") and c["sample"]
    ck = d["clocks"]
    assert ck["sm_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]  # the end-to-end number is not a copy of the device-resident one


def test_reference_arm_line():
    d = _line("r01_final_bench_reference.json")
    _common(d)
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"]


@pytest.mark.parametrize("name,n", [("r01_final2_bench_n2.json", 2), ("r01_final_bench_n8.json", 8)])
def test_multi_gpu_lines(name, n):
    d = _line(name)
    _common(d)
    assert d["n_gpus"] == n and f"x{n}" in d["config"]["parallelism"]
    one = _line("r01_final_bench.json")
    assert d["value"] > 0.8 * n * one["value"]  # weak scaling: whole-job aggregate


# ---- round 2 lines ---------------------------------------------------------------------------------------------
def test_r02_b200_arm_line():
    d = _line("r02_final_bench.json")
    _common(d)
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["parity_checked"] is True
    r = d["roofline"]
    assert r["kernel"] == "k_gather" and r["bound"] == "hbm" and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] == "port" and set(c["legs_s"]) >= {"scan_zero_copy_s", "fanout_alloc_s"}
    assert c["value"] >= c["value_faithful"]  # the headline CPU number is the FASTEST variant of every leg
    lat = d["latency"]
    assert lat["parity_checked"] and lat["device_us"] > 0 and lat["e2e_us"] > lat["device_us"]
    comp = d["extra"]["compaction"]
    assert comp["parity_checked"] and comp["records"] in (100_000_000, 10_000_000) and comp["roofline"]["frac"] > 0
    assert d["fanout_alone"]["us_per_burst"] > 0 and d["extra"]["write_path"]["ops_per_s"] > 0
    assert d["e2e"]["value"] < d["value"]


def test_r02_reference_arm_line():
    d = _line("r02_final_bench_reference.json")
    _common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    legs = d["cpu_baseline"]["legs_s"]
    fastest = min(legs["scan_faithful_s"], legs["scan_zero_copy_s"]) + legs["lists_s"] + \
        min(legs["fanout_alloc_s"], legs["fanout_prealloc_s"])
    assert abs(d["ms_per_step"] / 1e3 - fastest) < 1e-6


@pytest.mark.parametrize("name,n,mode", [("r02_bench_n2.json", 2, "weak"), ("r02_bench_n8.json", 8, "weak"),
                                         ("r02_bench_strong_n8.json", 8, "strong")])
def test_r02_multi_gpu_lines(name, n, mode):
    d = _line(name)
    _common(d)
    assert d["n_gpus"] == n and d["scaling"] == mode and d["parity_checked"] is True
    assert d["parity"]["merged_list_kvs"] > 0  # the broad List merged across shards was compared with the unsharded oracle
    assert set(d["cursor_exchange_us"]) >= {"nccl"}
