"""Host-side logic of bench.py that needs no GPU: the clock sampler's bookkeeping (which samples fell inside the timed region, the load kept
running until enough samples arrived, a fixed number of extra steps when several ranks step together) and the byte formulas of the rooflines."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_clock_sampler_summary_counts_the_samples_inside_the_timed_region():
    c = bench.ClockSampler(0)
    assert c.summary(0.0, 1.0)["samples"] == 0
    c.rows = [(0.5, 1965.0, 1965.0, []), (1.5, 1950.0, 1965.0, ["sw_power_cap"]), (2.5, 1965.0, 1965.0, [])]
    s = c.summary(1.0, 2.0)
    assert s["samples"] == 3 and s["samples_inside_timed_region"] == 1
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"]


def test_clock_sampler_keeps_the_load_running_until_enough_samples():
    class Alive:
        def poll(self):
            return None
    c = bench.ClockSampler(0)
    c.proc = Alive()
    n = []

    def step():
        n.append(1)
        if len(n) % 3 == 0:
            c.rows.append((time.perf_counter(), 1965.0, 1965.0, []))
    c.keep_load(step, want=3, limit_s=5.0)
    assert len(c.rows) == 3 and len(n) == 9 and c.extra_steps == 9
    # several ranks: exactly the number of steps rank 0 decided, whatever the samples say
    n.clear()
    c.keep_load(step, fixed=4)
    assert len(n) == 4 and c.extra_steps == 4
    # no child process (nvidia-smi missing): no extra steps, no hang
    c2 = bench.ClockSampler(0)
    c2.keep_load(step)
    assert c2.extra_steps == 0


def test_visible_device_list_maps_the_local_rank_to_the_physical_index(monkeypatch):
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6")
    assert bench.ClockSampler(1).index == 5
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-abcdef")
    assert bench.ClockSampler(0).index == 0


def test_hb_build_bytes_matches_the_formula_in_design():
    # reads E_b(8+2S) + E_o(12+9S) + 4S NP + 2S NL, writes 6S E_b + 9S n_off + 6S NP + 3S NL + S N  (198 MB at synth-2M, FP64)
    NP, NL, Eb, Eo, S = 200000, 50000, 1988231, 199999, 8
    n_off = Eo
    N = 3 * NP + 2 * NL
    want = Eb * (8 + 2 * S) + Eo * (12 + 9 * S) + 4 * S * NP + 2 * S * NL + 6 * S * Eb + 9 * S * n_off + 6 * S * NP + 3 * S * NL + S * N
    got = bench.hb_build_bytes(NP, NL, Eb, Eo, n_off, S)
    assert got == want and abs(got - 197952476) < 1e6
