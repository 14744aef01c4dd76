"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU path, oracle/_ref or the oracle port) prints the
one JSON line the driver parses, for both workloads, on a small grid; the argument surface of the other arm is what DESIGN.md §6 says."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run("--impl", "reference", "--grid", "24", "--steps", "5", "--warmup", "3")
    assert d["impl"] == "reference" and d["metric"] == "cg_iters_per_s" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["same_config"] is True
    assert d["steps_timed"] == 5                                   # real iterations on the named matrix, nothing extrapolated
    assert d["config"]["rows"] == 24 ** 3 and d["config"]["nnz"] == 7 * 24 ** 3 - 6 * 24 ** 2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_gmres_workload():
    d = _run("--impl", "reference", "--workload", "gmres", "--grid", "16", "--steps", "50")
    assert d["metric"] == "gmres_iters_per_s" and d["steps_timed"] == 51 and d["value"] > 0      # one restart cycle: 1 + 50 operator applications


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--grid", "16"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_does_not_load_the_product():
    """the arm times the reference: it must not map libhalab200.so (the judge looks at the loaded libraries)"""
    text = open(os.path.join(ROOT, "bench.py")).read()
    body = text[text.index("def run_reference"):text.index("# ---------------------------------------------------------------------------------------------------- the reference's GPU path")]
    assert "hala_b200" not in body.replace("hala_b200/matgen", "")
