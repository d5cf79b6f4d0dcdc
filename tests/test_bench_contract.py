"""The driver's contract with bench.py, as far as it can be held without a GPU: the reference arm (`--impl reference`: the
CPU oracle timed on the host cores) prints exactly one JSON line on stdout with the keys the driver reads, on the same
workload description as the GPU arm, and the `--models` knob shows in it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                          "--warmup", "1", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # stdout carries the result line and nothing else
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    d = _run()
    assert d["impl"] == "reference" and d["metric"] == "task_x_worker_cost_evaluations_per_sec" and d["unit"] == "evals/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"].startswith("synthetic") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg1") and d["config"]["n_asks"] == 1000 and d["config"]["n_workers"] == 10000
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == cb["host_cpus"]["threads"] <= cb["host_cpus"]["affinity"]      # the threads it really ran on
    assert d["gpu_launches"] == 0


def test_models_knob_shows_in_the_workload():
    d = _run("--models", "200")
    assert "200 distinct worker model strings" in d["config"]["workload"]
