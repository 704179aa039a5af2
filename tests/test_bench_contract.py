"""bench.py's JSON contract on the arm that runs without a GPU (--impl reference: the reference's own
fluid.cpp, or the C port when oracle/_ref is absent), and the workload description helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--width", "512", "--height", "512", "--iters", "30"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "cell-updates/s per pressure iter" and d["unit"] == "cell-updates/s" and d["dtype"] == "f32"
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 1e6 and d["config"]["workload"]


def test_non_zero_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_shapes():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(gpus=1, width=0, height=0, iters=100)
    assert bench.workload_shape(a) == (4096, 4096, 100)
    a = argparse.Namespace(gpus=8, width=0, height=0, iters=100)
    assert bench.workload_shape(a) == (16384, 16384, 100)
    assert "16384" in bench.workload_config(a)["workload"]
