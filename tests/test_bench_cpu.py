"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys, running the unmodified
reference from oracle/_ref (or the oracle port when it is absent) on a bounded sample -- here a tiny one."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--particles", "20", "--ref-batch", "16",
                          "--sweeps", "2", "--steps", "2", "--warmup", "1", "--no-extras", *extra],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run()
    assert d["impl"] == "reference" and d["metric"] == "particle-sweeps/sec" and d["unit"] == "particle-sweeps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["config"]["S"] == 20 and d["config"]["ref_batch"] == 16 and "workload" in d["config"]


def test_reference_arm_shrinks_its_sample_to_the_cpu_budget():
    d = _run("--ref-seconds", "0.000001")
    assert d["config"]["ref_batch"] == 8                      # the floor: one multiple of 8
    assert "B=8" in d["cpu_baseline"]["sample"]


def test_reference_arm_on_other_ranks_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
