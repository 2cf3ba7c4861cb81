"""bench.py host-side pieces that can be checked without a GPU: the clock sampler
(against a fake `nvidia-smi`), the reference arm's one-line contract."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FAKE = """#!/bin/bash
sleep 0.05
i=0
while true; do
  if [ $i -ge 12 ] && [ $i -le 14 ]; then
    echo "0, 1500, 1965, 990.0, 0x4, Not Active, Not Active, Not Active, Active"
  else
    echo "0, $((1900 + i)), 1965, 600.0, 0x0, Not Active, Not Active, Not Active, Not Active"
  fi
  i=$((i+1)); sleep 0.02
done
"""


def test_clock_sampler_windows(tmp_path, monkeypatch):
    import bench
    fake = tmp_path / "nvidia-smi"
    fake.write_text(FAKE)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0, period_ms=20)
    s.start()
    time.sleep(0.22)                                   # warm-up: process already streaming
    t0 = time.time()
    time.sleep(0.30)                                   # the "timed region"
    t1 = time.time()
    time.sleep(0.05)
    r = s.stop(t0, t1)
    assert r["window"] == "timed region" and 8 <= r["samples"] < r["samples_total"]
    assert r["sm_max_mhz"] == 1965.0 and 1500 <= r["sm_mhz"] <= 1965 and r["period_ms"] == 20
    assert r["reasons"] == ["sw_power_cap"] and r["power_w_max"] == 990.0
    # a timed region shorter than the sampling period: nearest samples under the same load
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.2)
    t0 = time.time()
    r = s.stop(t0, t0 + 1e-4)
    assert r["samples"] >= 1 and "warm-up" in r["window"] and r["sm_mhz"] is not None
    # no nvidia-smi on the machine
    monkeypatch.setenv("PATH", str(tmp_path / "nothing"))
    s = bench.ClockSampler(0)
    s.start()
    assert s.stop(0, 1)["reasons"] == ["nvidia-smi unavailable"]


def test_reference_arm_prints_one_json_line_per_job():
    env = dict(os.environ, RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    rec = json.loads(r.stdout.strip())
    assert rec["impl"] == "reference" and rec["unavailable"]
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
