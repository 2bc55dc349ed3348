"""bench.py's output contract, checked on the committed bench lines (profiles/*.json are the driver-format
lines of real B200 runs) and on the argument parser -- no GPU needed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}


def _line(name):
    for raw in open(os.path.join(ROOT, "profiles", name)):
        if raw.startswith("{"):
            return json.loads(raw)
    raise AssertionError(name)


def test_native_line_carries_the_contract_keys():
    d = _line("r2k_bench_default.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] in base["metric"] and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_final_line_carries_the_segmentation_extras():
    """the last bench line of the round: the contract keys plus extra.segmentation"""
    d = _line("r2u_bench_default.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    sg = d["extra"]["segmentation"]
    assert abs(sg["ms_per_field"] - (sg["normalize_ms"] + sg["unet_ms"] + sg["instances_ms"])) < 1e-6
    assert sg["instances"] > 0 and sg["candidates"] > sg["instances"] and sg["gpu_launches_per_field"] > 0
    assert abs(sg["unet_tflops"] - sg["unet_gflop"] / sg["unet_ms"]) < 1e-6 * sg["unet_tflops"] + 1e-9
    assert sg["cpu_port"]["kind"] == "port" and sg["cpu_port"]["labels_equal_gpu"] is True
    ch = sg["chain"]
    assert ch["scored_cells"] > 0 and abs(ch["cells_per_s"] - ch["scored_cells"] * 1e3 / ch["ms_per_field"]) < 1e-3


def test_reference_line():
    d = _line("r2k_bench_reference.json")
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_multi_gpu_line_is_weak_scaling_with_config5_strains():
    d = _line("r2f_bench_2gpu.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["strains"] == 100


def test_cli_defaults():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
