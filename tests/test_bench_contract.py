"""bench.py's JSON contract, checked without a GPU: the reference arm runs live on a small file set (it needs no device), and the committed
lines of the CUDA arm (profiles/r02_bench_n*.json, produced on B200 boxes) must carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e")


def test_reference_arm_runs_on_cpu_and_prints_one_contract_line(tmp_path):
    d = str(tmp_path / "files")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--files", "6", "--steps", "2", "--warmup", "1", "--dir", d],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in BASE:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "GiB/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    cfg = line["config"]
    assert "workload" in cfg and cfg["errors"] == 0
    assert set(cfg["configurations_one_pass_gib_s"]) >= {"readme_8_workers_sqpoll"} and cfg["timed_configuration"] in cfg["configurations_one_pass_gib_s"]
    assert cfg["object_store_like_gib_s"] > 0


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_cuda_arm_lines_carry_the_contract(n):
    path = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    line = json.loads(open(path).read())
    for k in BASE + ("gpu_launches", "clocks", "roofline", "io_roofline", "parity"):
        assert k in line, k
    assert line["n_gpus"] == n and line["scaling"] == "weak" and line["dtype"] == "u8" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["metric"] == "GiB/s of byte ranges landed in HBM" and "workload" in line["config"] and line["warmup"] >= 3
    e = line["e2e"]
    assert e["unit"] == "GiB/s" and e["h2d_bytes_per_step"] == 20_578_304_000 and e["d2h_bytes_per_step"] == 4 * 78_500 and e["bad_chunks"] == 0
    # `value` is the path itself now: it and the Python-mirror figure describe the same thing
    assert 0.85 < e["value"] / line["value"] < 1.1
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and 0.9 < r["frac"] < 1.05 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert line["gpu_launches"] > 0 and line["clocks"]["sm_mhz"] and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    par = line["parity"]
    assert par["expected_crc_source"] == "oracle" and par["bad_chunks_all_ranks"] == 0 and par["landed_bytes_vs_file_bytes"]["mismatched_bytes_all_ranks"] == 0
    if n == 1:
        cb = line["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] == 16 and "sample" in cb and cb["value"] > 0
    else:
        assert line["c4"]["bad_chunks"] == 0 and line["two_devices_one_process"]["bad_chunks"] == 0
    if n == 8:
        assert line["c5"]["bad_chunks"] == 0 and line["c5"]["bytes_landed_all_ranks"] > 0.9e12
