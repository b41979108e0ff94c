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
    from oracle import bench_cpu
    n = bench_cpu.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    assert bench.WORKLOADS["portal_in_portal"] == (3840, 2160, 40)          # BASELINE.json headline config
    assert bench.WORKLOADS["mobius_monoportal"] == (7680, 4320, 64) and bench.WORKLOADS["basics"] == (256, 256, 4)


def test_both_arms_name_the_workload_alike_and_pins_exist():
    """The driver compares `config.workload` of the two arms: one function makes both.  Every BASELINE config has full-size
    sha256 pins (float and RGBA8 frames) for bench.py's `parity` block; config 5 also for orbit frames 0 and 180."""
    import argparse
    for scene, (w, h, d) in bench.WORKLOADS.items():
        a = argparse.Namespace(scene=scene, width=w, height=h, depth=d, orbit=0)
        assert bench.workload_name(a) == f"{scene}.ron {w}x{h} depth {d}, saved camera, aa 1"
        pins = bench.golden_pins(scene, w, h, d)
        assert pins and len(pins["sha256_f32_rgba"]) == 64 and len(pins["sha256_rgba8"]) == 64
    for k in (0, 180):
        assert len(bench.golden_pins("mobius_monoportal", 7680, 4320, 64, 360, k)["sha256_rgba8"]) == 64
    assert bench.golden_pins("mobius_monoportal", 7680, 4320, 64, 360, 7) is None
    assert bench.golden_pins("portal_in_portal", 640, 360, 40) is None


def test_cpu_baseline_runs_in_a_clean_subprocess_whatever_the_parent_environment():
    """oracle/bench_cpu.py resets affinity and the OpenMP environment (a torchrun child has OMP_NUM_THREADS=1): the thread count
    an OpenMP region really gets is what bench.py reports as `cores`."""
    import argparse
    import json
    import subprocess
    a = argparse.Namespace(scene="basics", width=256, height=256, depth=4, orbit=0)
    old = dict(os.environ)
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        res = bench.cpu_oracle(a, budget_s=0.3, steps=3)
    finally:
        os.environ.clear()
        os.environ.update(old)
    from oracle import bench_cpu
    assert res["threads_used"] == bench_cpu.usable_cores() and res["steps"] == 3 and res["value"] > 0
    one = bench.cpu_oracle(a, budget_s=0.2, threads=1)
    assert one["threads_used"] == 1
    # the reference arm prints the line the driver expects, with the same workload string
    env = dict(os.environ, RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scene", "basics", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"]["workload"] == bench.workload_name(a) and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
