"""CPU: the reference arm of bench.py (`--impl reference`) - the one leg of the benchmark that needs no GPU - on a
tiny configuration: one JSON line with the contract's keys, rank 0 only under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
       "--rows", "40000", "--queries", "64", "--ref-blocks", "2", "--ref-block-rows", "5000"]


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run(CMD, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_one_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and res.stdout.strip() == lines[0]          # stdout carries the line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"] == "queries/sec, exact top-100, 25.7Mx768" and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["extrapolated"] is True
    assert d["ms_per_step"] == d["sample_ms_per_step"]                 # the MEASURED sample, not the projection
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "blocks" in cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["rows"] == 40000 and d["config"]["queries"] == 64 and d["config"]["sample_rows_per_step"] == 10000
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_0_only():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == "", (res.stdout[-500:], res.stderr[-500:])
