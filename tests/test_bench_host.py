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


def test_ncu_traffic_reads_the_committed_capture():
    """roofline.traffic comes from profiles/r02_kernels_10m_ncu_full.txt: DRAM read + write bytes of one launch"""
    t = bench.ncu_traffic("k_decode_vertex_scan")
    assert t is not None and 0.5e9 < t < 1.5e9          # 0.80 GB read + 0.08 GB written at 10M vertices
    f = bench.ncu_traffic("k_flatten_halfedges")
    assert f is not None and 1.5e9 < f < 2.0e9          # 12 B in, 16 B out per half-edge, 60M half-edges
    assert bench.ncu_traffic("(k_encode_vtx_packed<T, NC>)") is not None
    assert bench.ncu_traffic("k_no_such_kernel") is None


def test_bind_rank_without_a_gpu_keeps_the_affinity():
    """no CUDA device (this container) or an unknown topology: the rank stays where it is"""
    import os
    from harry_b200 import shard
    before = os.sched_getaffinity(0)
    assert shard.bind_rank_to_gpu_node(0, "/nonexistent") == before
    assert os.sched_getaffinity(0) == before
