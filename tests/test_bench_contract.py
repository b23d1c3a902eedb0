"""CPU: bench.py's reference arm runs here (no GPU), prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "64", "--ref-sample", "16"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    # the unmodified reference under multiprocessing.Pool where a copy of it exists (build container: /root/reference,
    # GPU box: oracle/_ref installed by build()); the C port of it otherwise
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refbench
    assert d["cpu_baseline"]["kind"] == ("reference" if refbench.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > (5e2 if d["cpu_baseline"]["kind"] == "reference" else 1e4)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_flop_accounting_matches_design():
    sys.path.insert(0, ROOT)
    import bench
    # one accepted step of one row: 12 RHS x 38 + 6 x 158 + 20 for the step, + HINIT (1 RHS + 60)
    assert bench.algorithmic_flops(1, 1, 1) == 12 * 38 + 6 * 158 + 20 + 38 + 60
    # a rejected attempt costs 11 RHS and the stage arithmetic
    assert bench.algorithmic_flops(2, 1, 1) - bench.algorithmic_flops(1, 1, 1) == 11 * 38 + 6 * 158 + 20
