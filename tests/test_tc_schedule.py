"""Static tile schedule of the tcgen05 filter (scema_b200/csrc/tc_sched.h), enumerated on the host: every tile of the
upper triangle inside the launch's row / column range is visited exactly once over all shards and units, for both
kernel flavours (one CTA per 128-row tile, CTA pair per 256-row tile), incl. the column panels of the host-buffer
pipeline and the row panels of the streamed compare."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_schedule_covers_every_tile_once(tmp_path):
    exe = str(tmp_path / "tc_sched_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "helpers", "tc_sched_check.cc")])
    r = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:]
