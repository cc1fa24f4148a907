"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the clock sampler's parsing."""
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ngp_lego_train_rays_per_s" and line["unit"] == "rays/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_selects_the_timed_window():
    import bench
    cs = bench.ClockSampler(0)
    cs.proc = type("P", (), {"terminate": lambda self: None})()
    cs.t = type("T", (), {"join": lambda self, timeout=None: None})()
    t0 = datetime.datetime(2026, 1, 2, 3, 4, 5).timestamp()

    def row(dt, sm, cap="Not Active"):
        ts = datetime.datetime.fromtimestamp(t0 + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return f"{ts}, {sm}, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, {cap}\n"

    cs.lines = [row(-1.0, 300), row(-0.5, 1200)] + [row(0.01 * k, 1950 + k) for k in range(1, 8)] + [row(2.0, 400, "Active")]
    out = cs.stop(t0, t0 + 0.1)
    assert out["window"] == "timed region" and out["samples"] == 7 and out["sm_mhz"] == 1954 and out["sm_max_mhz"] == 1965 and out["reasons"] == []
    cs.lines = [row(-1.0, 1900, "Active"), row(5.0, 1910)]
    out = cs.stop(t0, t0 + 0.1)                                  # too few samples inside: falls back to everything under load
    assert out["window"] != "timed region" and out["samples"] == 2 and out["reasons"] == ["sw_power_cap"]
