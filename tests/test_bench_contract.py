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


def test_clock_sampler_summary_and_bracketing(monkeypatch):
    """The clocks object of the bench line: median SM clock and throttle reasons of the samples that fall inside the timed
    region; a region without samples reports the last one before it; short regions are bracketed (no polling), and a
    missing nvidia-smi yields the 'unavailable' record instead of an exception."""
    sys.path.insert(0, ROOT)
    import bench
    import time as _time
    s = bench.ClockSampler(0)
    row = "0, {sm}, 1965, 990.5, 0x4, Not Active, Not Active, Not Active, {cap}, 3996"
    t0 = _time.perf_counter()
    s.lines = [(t0 - 2.0, row.format(sm=1965, cap="Not Active")), (t0 + 0.1, row.format(sm=1500, cap="Active")),
               (t0 + 0.3, row.format(sm=1600, cap="Active"))]
    s.begin = t0
    rec = s._summary({})
    assert rec["sm_mhz"] == 1550.0 and rec["sm_max_mhz"] == 1965.0 and rec["reasons"] == ["sw_power_cap"] and rec["samples"] == 2
    assert rec["mem_mhz"] == 3996.0 and abs(rec["power_w"] - 990.5) < 1e-9
    s.begin = t0 + 10.0                                   # nothing inside the region: the nearest earlier sample stands in
    rec = s._summary({})
    assert rec["samples"] == 0 and rec["sm_mhz"] == 1600.0 and "note" in rec

    calls = []
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: calls.append("start"))
    monkeypatch.setattr(bench.ClockSampler, "snapshot", lambda self: calls.append("snapshot") or setattr(self, "bracket", True))
    monkeypatch.setattr(bench.time, "sleep", lambda x: None)
    short = bench.ClockSampler(0)
    short.begin_region(0.05)
    assert calls == ["snapshot"] and getattr(short, "after_snapshot", False)
    long_ = bench.ClockSampler(0)
    long_.begin_region(20.0)
    assert calls == ["snapshot", "start"] and not getattr(long_, "after_snapshot", False)
    # no nvidia-smi on this machine (or none was started): the record says so
    assert bench.ClockSampler(0).end_region()["reasons"] == ["nvidia-smi unavailable"]
