"""CPU: bench.py's reference arm honours the driver contract - exactly ONE JSON line on stdout (library banners and
progress go to stderr), the metric / unit / config of BASELINE.json, a cpu_baseline describing the run, and an e2e
object.  (The GPU arm prints the same line plus roofline / clocks / gpu_launches; it needs a B200.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "hmsg_rgbd_frames_per_s_ingested" and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["data"] == "synthetic" and j["n_gpus"] == 1
    assert j["steps"] == 1 and j["warmup"] == 0 and j["value"] > 0 and j["ms_per_step"] > 0
    assert "640x480" in j["config"]["workload"] and "configs[1]" in j["config"]["workload"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["unit"] == "frames/s" and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
