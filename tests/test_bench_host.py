"""CPU: host-side logic of bench.py that needs no GPU (the clock sampler's choice of samples)."""
import time

import bench


def _line(sm, reason_power="Not Active"):
    return f"0, {sm}, 1965, 300.0, 0x0, Not Active, Not Active, Not Active, {reason_power}"


def test_clock_sampler_reports_samples_inside_the_timed_region():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    now = time.perf_counter()
    s.lines = [(now - 1.0, _line(1200)), (now - 0.5, _line(1300)), (now + 0.001, _line(1965, "Active")), (now + 0.002, _line(1950))]
    s.t0 = now
    time.sleep(0.01)
    out = s.stop()
    assert out["samples"] == 2 and out["window"] == "timed region"
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_falls_back_to_the_last_samples():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    now = time.perf_counter()
    s.lines = [(now - 1.0, _line(1200)), (now - 0.5, _line(1900)), (now - 0.4, _line(1965)), (now - 0.3, _line(1965))]
    s.t0 = now
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["window"].startswith("last samples")


def test_clock_sampler_without_nvidia_smi():
    out = bench.ClockSampler(0).stop()
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]
