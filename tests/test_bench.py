"""CPU tests of bench.py's own plumbing (no GPU): the clock sampler that watches the timed regions, and the reference arm's JSON line."""
import importlib.util
import json
import os
import stat
import subprocess
import sys
import time

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_clock_sampler_runs_once_and_keeps_only_samples_from_the_timed_regions(tmp_path, monkeypatch):
    """One nvidia-smi for the whole run, started (and through its start-up) BEFORE the first timed region; only lines that arrive
    inside a region are summarised; throttle reasons are reported by name; a machine without nvidia-smi gives an empty summary."""
    bench = _bench()
    fake = tmp_path / "nvidia-smi"
    log = tmp_path / "starts.log"
    fake.write_text("#!/bin/bash\necho started >> %s\nsleep 0.3\nwhile true; do echo '0, 1965, 1965, 500.0, 0x0, Not Active, Not Active, Not Active, Active'; sleep 0.05; done\n" % log)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    c = bench.ClockSampler(0)
    t0 = time.perf_counter()
    c.start()
    assert time.perf_counter() - t0 >= 0.25 and len(c.rows) >= 1  # start() returns once the first line is out: start-up is over
    outside = len(c.rows)
    time.sleep(0.2)  # not a timed region: these lines must not count
    with c:
        time.sleep(0.02)  # shorter than the sampling period: still gets its sample
    with c:
        time.sleep(0.25)
    s = c.summary()
    assert log.read_text().count("started") == 1  # re-entering did not start another process
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"]
    assert 3 <= s["samples"] < len(c.rows) - outside + 1 and c.proc is None
    assert c.summary() == s  # idempotent after the process is gone

    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    none = bench.ClockSampler(0)
    with none:
        pass
    assert none.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_reference_arm_prints_the_contract_line_without_a_gpu(oracle_mod):
    """bench.py --impl reference on a bounded sample (every 16th pixel of the full-size workload): one JSON line with the contract's
    keys, every host thread, no device work."""
    if not (oracle_mod.available("reference") or oracle_mod.available("port")):
        import pytest

        pytest.skip("no CPU checker built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-stride", "16"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "Mrays/s" and line["n_gpus"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["config"]["workload"].startswith("config3")
