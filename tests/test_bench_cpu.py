"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference's dequant -> fp16 matmul path)
prints ONE JSON line with the keys the driver reads, on a tiny model so that it runs in seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                       env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run("--impl", "reference", "--model", "tiny", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    e = {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--model", "tiny"],
                       capture_output=True, text=True, cwd=ROOT, env={**os.environ, **e}, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
