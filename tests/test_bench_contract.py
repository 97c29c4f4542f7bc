"""bench.py's contract on a machine without a GPU: the reference arm prints ONE JSON line with the agreed keys (rank 0 only
under torchrun), and the product arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT,
                          timeout=600)


def test_reference_arm_line():
    r = _bench(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("ORB frames/s @640x480/1000kp") and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", r.stdout + r.stderr


def test_product_arm_fails_loudly_without_a_gpu():
    import vo_slam_test_b200 as v
    if v.device_count() > 0:
        import pytest
        pytest.skip("GPU present")
    r = _bench(["--steps", "1", "--warmup", "1", "--frames", "8", "--skip-map", "--skip-cpu", "--skip-single"])
    assert r.returncode != 0 and r.stdout.strip() == "", r.stdout
