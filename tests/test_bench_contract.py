"""Host-side pieces of bench.py that can be checked without a GPU."""
import datetime
import os
import sys
import time

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _Proc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        pass


def _line(t, sm, power_cap="Not Active"):
    ts = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    return f"{ts}, 0, {sm}, 1965, 400.0, 0x0, Not Active, Not Active, Not Active, {power_cap}\n"


def _sampler(lines):
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.f = open(s.path, "w")
    s.f.write("".join(lines))
    s.f.flush()
    return s


def test_clock_sampler_windows():
    now = time.time()
    # enough samples inside the timed region: only those count (the idle 1200 MHz sample before it does not)
    s = _sampler([_line(now - 3, 1200), _line(now - 0.9, 1965), _line(now - 0.7, 1950, "Active"), _line(now - 0.5, 1965)])
    c = s.stop(now - 1.0, now - 0.4, now)
    assert c == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3, "window": "timed region"}
    # a timed region shorter than the sampling period: the untimed repeat of the same steps supplies the samples
    s = _sampler([_line(now - 3, 1200), _line(now - 0.5, 1965), _line(now - 0.3, 1965), _line(now - 0.1, 1950)])
    c = s.stop(now - 1.0, now - 0.98, now)
    assert c["samples"] == 3 and c["sm_mhz"] == 1965.0 and "untimed repeat" in c["window"]
    # nothing usable at all: report what there is, never crash
    s = _sampler(["garbage\n"])
    c = s.stop(now - 1.0, now - 0.98, now)
    assert c["sm_mhz"] is None and c["samples"] == 0
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]


def test_usable_cores_and_workloads():
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    assert bench.WORKLOADS["portal_in_portal"] == (3840, 2160, 40)          # BASELINE.json headline config
    assert bench.WORKLOADS["mobius_monoportal"] == (7680, 4320, 64) and bench.WORKLOADS["basics"] == (256, 256, 4)
