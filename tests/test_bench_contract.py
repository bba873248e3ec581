"""CPU tests of the bench contract: the committed bench lines carry every key the driver reads, and the
`--impl reference` arm (the oracle's AVX2 restatement on the host cores) runs end to end without a GPU."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"]


def _lines():
  return sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_c2_final*.json")) +
                glob.glob(os.path.join(ROOT, "profiles", "r01_bench_c5_100m_1gpu_v3.json")) +
                glob.glob(os.path.join(ROOT, "profiles", "r01_bench_c4_10m_1gpu.json")))


@pytest.mark.parametrize("path", _lines(), ids=[os.path.basename(p) for p in _lines()])
def test_committed_bench_lines_follow_the_contract(path):
  d = json.load(open(path))
  for k in BASE_KEYS + ["gpu_launches", "clocks", "roofline"]:
    assert k in d, k
  assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
  assert "workload" in d["config"] and "model" not in d["config"]
  for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
    assert k in d["e2e"], k
  assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
  assert d["e2e"]["value"] != d["value"]                       # measured separately, not a copy of the device number
  r = d["roofline"]
  for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
    assert k in r, k
  assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
  assert d["gpu_launches"] > 0
  assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
  if "cpu_baseline" in d:
    for k in ("value", "unit", "cores", "kind", "sample"):
      assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("port", "reference")


def test_reference_arm_runs_on_the_host_cores():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1_synthetic",
                        "--steps", "1", "--warmup", "0", "--cpu-sample", "200"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, out.stdout                             # exactly one JSON line on stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["value"] > 0
  for k in BASE_KEYS + ["cpu_baseline"]:
    assert k in d, k
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
