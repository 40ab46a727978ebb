"""bench.py on the CPU: the reference arm's JSON line (contract keys) and the clock sampler's parsing."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["unit"] == "cells/s" and d["value"] > 0 and d["metric"].startswith("gridcell-fits/sec")
    sys.path.insert(0, ROOT)
    import bench
    from oracle import reference_runner

    # the unmodified reference when it is installed (baseline/_ref travels to the GPU box), else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_runner.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"] == bench.bench_config()            # identical to the CUDA arm's config (driver: same_config)


def test_reference_arm_sample_is_bounded():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.reference_cells_per_step(1, 16) == 16 and bench.reference_cells_per_step(16, 16) == 16
    assert bench.reference_cells_per_step(40, 16) == 6 and bench.reference_cells_per_step(40, 32) == 12
    assert bench.reference_cells_per_step(10_000, 16) == 1
    for steps in (1, 3, 20, 40, 100, 256):
        for cores in (8, 16, 32, 64):
            assert steps * bench.reference_cells_per_step(steps, cores) <= max(16 * cores, steps)


def test_clock_sampler_keeps_the_samples_of_the_timed_region(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    # a stand-in nvidia-smi that prints one line every 20 ms with the current time stamp
    fake = tmp_path / "nvidia-smi"
    fake.write_text(
        "#!" + sys.executable + "\n"
        "import sys, time, datetime\n"
        "i = 0\n"
        "while True:\n"
        "    ts = datetime.datetime.now().strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]\n"
        "    clk = 1965 if i % 7 else 1200\n"
        "    print(f'{ts}, 0, {clk}, 1965, 400.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active', flush=True)\n"
        "    i += 1\n"
        "    time.sleep(0.02)\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.4)
    s.begin_region()
    time.sleep(0.3)
    s.end_region()
    time.sleep(0.1)
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0 and out["sm_mhz"] == 1965.0          # median of the busy samples
    assert out["reasons"] == ["sw_power_cap"]
    assert 5 <= out["samples"] <= 25 and "note" not in out                    # only the ~15 samples of the 0.3 s region
    # a region shorter than the sampling period falls back to the samples since the start, and says so
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.3)
    s.begin_region()
    s.end_region()
    out = s.stop()
    assert out["sm_mhz"] is not None and ("note" in out or out["samples"] >= 1)
