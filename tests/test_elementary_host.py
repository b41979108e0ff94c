"""The elementary functions the numeric profile pins -- sin cos tan asin acos atan exp2 log2 and the
exp / log / pow derived from them: the device header's definitions against the oracle's, bit for bit, on the host, plus
their accuracy against double-precision libm (tests/host_harness/elementary_check.cpp)."""
import os
import subprocess

from conftest import ROOT


def test_device_and_oracle_elementary_functions_agree_bitwise(tmp_path):
    exe = str(tmp_path / "elementary_check")
    cc = subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "host_harness"),
                         "-I", os.path.join(ROOT, "portal_b200", "csrc", "device"), "-I", os.path.join(ROOT, "oracle"),
                         os.path.join(ROOT, "tests", "host_harness", "elementary_check.cpp"), "-o", exe], capture_output=True, text=True, timeout=600)
    assert cc.returncode == 0, cc.stderr[-3000:]
    run = subprocess.run([exe, "1500000"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:]
    w = run.stdout.split()
    stats = {w[i]: float(w[i + 1]) for i in range(0, len(w), 2)}
    assert stats["mismatches"] == 0 and stats["cases"] >= 1_500_000
    # exp2 / log2 / exp are ~1 ulp Cephes kernels, log = log2 * ln 2; pow inherits exp2(y * log2 x)'s conditioning (GLSL ES 3.00, 4.5.1)
    assert stats["exp2_rel"] < 1.5e-7 and stats["log2_rel"] < 2.5e-7 and stats["log_rel"] < 3e-7 and stats["exp_rel"] < 2e-7
    assert stats["pow_rel"] < 2e-5
