"""CPU checks of bench.py's host-side helpers (no GPU, no timing)."""
import importlib.util
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("latte_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Done:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass


def _sampler(bench, lines, t_load, t_mark):
    s = bench.ClockSampler(0)
    s.proc = _Done()
    s.sm_max = 1965.0
    s.lines = lines
    s.t_load, s.t_mark = t_load, t_mark
    return s


def test_clock_sampler_reports_the_samples_of_the_timed_region():
    """Two NVML fields per sample: SM clock and the event-reason bitmask (0x4 = sw_power_cap,
    0x8 = hw_slowdown, 0x20 / 0x40 = thermal slowdowns)."""
    bench = _bench()
    now = time.time()
    lines = [(now - 1.00, "2026/10/18 00:00:00.000, 120, 0x0000000000000001"),      # idle, before the load
             (now - 0.30, "2026/10/18 00:00:00.700, 1965, 0x0000000000000000"),     # warm-up
             (now - 0.10, "2026/10/18 00:00:00.900, 1500, 0x0000000000000004"),     # timed region
             (now - 0.05, "2026/10/18 00:00:00.950, 1400, 0x0000000000000004")]
    out = _sampler(bench, lines, now - 0.4, now - 0.15).stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1450.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert out["window"].startswith("timed region")


def test_clock_sampler_falls_back_to_the_continuous_load_and_decodes_slowdowns():
    bench = _bench()
    now = time.time()
    lines = [(now - 2.00, "t, 120, 0x1"),
             (now - 0.30, "t, 1965, 0x0"),
             (now - 0.04, "t, 900, 0x48")]                                          # hw + hw-thermal slowdown
    out = _sampler(bench, lines, now - 0.5, now - 0.1).stop()
    assert out["samples"] == 2                        # one sample in the region: the load window is used
    assert "continuous load" in out["window"]
    assert out["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]
    # the idle sample before the load never enters the record
    assert out["sm_mhz"] == (1965 + 900) / 2


def test_bench_constants_name_the_headline_workload():
    bench = _bench()
    assert bench.N_GLOBAL == 32768 and bench.DIM == 512
    assert bench.METRIC == "clip_loss_fwd_bwd_samples_per_sec" and bench.UNIT == "samples/s"
    assert bench.CPU_BLOCK_ROWS == 4096
