"""Host-side logic of bench.py that needs no GPU: the clock sampler only counts lines that arrived inside the timed
region and falls back to sampling under the same load when the region was shorter than nvidia-smi's interval; the
measured-DRAM record in profiles/pgs_traffic.json is keyed on the solver sources it was taken on."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _FakeProc:
    def terminate(self): pass
    def wait(self, timeout=None): return 0
    def kill(self): pass


def _sampler():
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.lines, s.proc, s.t0, s.t1 = [], _FakeProc(), None, None
    return s


def test_only_lines_inside_the_timed_region_count():
    s = _sampler()
    s.lines.append((time.perf_counter(), "1200, 1965, Not Active, Not Active, Not Active, Not Active"))     # warm-up: ignored
    s.begin()
    s.lines.append((time.perf_counter(), "1965, 1965, Not Active, Not Active, Not Active, Active"))
    s.lines.append((time.perf_counter(), "1950, 1965, Not Active, Not Active, Not Active, Not Active"))
    s.end()
    s.lines.append((time.perf_counter() + 1.0, "300, 1965, Active, Not Active, Not Active, Not Active"))  # after: ignored
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and "note" not in out


def test_short_region_is_sampled_under_the_same_load_and_says_so():
    s = _sampler()
    s.begin(); s.end()                                     # nothing arrived in time
    calls = []

    def load():
        calls.append(1)
        s.lines.append((time.perf_counter(), "1965, 1965, Not Active, Not Active, Not Active, Not Active"))
    out = s.stop(load)
    assert len(calls) == 2 and out["samples"] == 2 and out["sm_mhz"] == 1965.0 and "note" in out


def test_traffic_record_matches_the_solver_sources():
    """A record taken on other kernel sources must not be used: bench.py compares this hash."""
    recs = json.load(open(os.path.join(ROOT, "profiles", "pgs_traffic.json")))
    mine = [r for r in recs if r.get("workload") == "marble_box_1k" and r.get("scenes") == 2664]
    assert mine, "no record for the default workload"
    assert mine[-1]["source_hash"] == bench.solver_source_hash(), "re-run tools/ncu_summary.py on a capture of the current kernels"
    for f in bench.SOLVER_SOURCES:
        assert os.path.exists(os.path.join(ROOT, "slrbs_b200", "csrc", f)), f
