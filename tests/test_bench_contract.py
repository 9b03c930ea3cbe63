"""bench.py's reference arm runs without a GPU: check the one-JSON-line contract on the smallest workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "graphene_40nm_f32_dos", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "nnz*moments*vectors/s"
    assert rec["value"] > 0 and rec["steps"] == 1 and rec["n_gpus"] == 1
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1 and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == dict(value=rec["value"], unit=rec["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in rec["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                          "graphene_40nm_f32_dos", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
