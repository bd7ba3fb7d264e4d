"""K1's reciprocal-based exact division (div_tab, scema_b200/csrc/resample.cu) equals IEEE division
for every divisor the spline tables contain — checked on the CPU with hardware FMA."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_div_tab_matches_ieee_division(tmp_path):
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA instruction")
    exe = str(tmp_path / "fastdiv_check")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe,
                           os.path.join(HERE, "helpers", "fastdiv_check.c"), "-lm"])
    r = subprocess.run([exe, "700", "200"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "bad(2-step)=0" in r.stdout
