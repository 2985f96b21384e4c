"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout with the keys the
driver reads, and the committed lines of our arm (profiles/r01_bench_*_final.json, plain runs on B200) carry the
roofline / cpu_baseline / e2e / clocks objects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_json_line():
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                     capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, r.stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and BASE_KEYS <= set(d)
  assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
  assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
  assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.parametrize("name", ["r01_bench_n1_final.json", "r01_bench_n2_final.json", "r01_bench_n8_final.json"])
def test_committed_bench_lines_follow_the_contract(name):
  d = json.loads(open(os.path.join(ROOT, "profiles", name)).read())
  assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
  assert d["gpu_launches"] == 2 * d["steps"] and d["scaling"] == "weak" and d["vs_baseline"] is None
  rf = d["roofline"]
  assert rf["bound"] in ("hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
  assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
  assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
  if d["n_gpus"] == 1:
    assert d["cpu_baseline"]["kind"] == "port" and 0.5 < min(v["frac"] for v in d["roofline_spline"].values()) < 1.0
