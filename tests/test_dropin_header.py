"""The drop-in header scema_b200/host/strain2spline_b200.h against the reference header itself: one
caller (tests/helpers/dropin_driver.cc, written only against the MatHistPredict API and shaped like
FE_problem.h:1091-1270) is compiled against both; stdout and every result file must be identical."""
import filecmp
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_driver_ref")


def build_ours(tmp_path):
    exe = str(tmp_path / "dropin_driver_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "scema_b200", "host"),
                           os.path.join(ROOT, "tests", "helpers", "dropin_driver.cc"), "-o", exe,
                           "-L" + os.path.join(ROOT, "scema_b200"), "-lscema_hist",
                           "-Wl,-rpath," + os.path.join(ROOT, "scema_b200")])
    return exe


@pytest.mark.parametrize("n_qp,n_steps,P,thr,seed", [(64, 12, 10, 1e-6, 5), (300, 21, 10, 4e-7, 9), (41, 7, 4, 1e-6, 2)])
def test_same_caller_same_bytes(tmp_path, n_qp, n_steps, P, thr, seed):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref not built")
    ours_exe = build_ours(tmp_path)
    outs = {}
    for name, exe in (("ref", REF_BIN), ("ours", ours_exe)):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, str(d), str(n_qp), str(n_steps), str(P), repr(thr), str(seed)], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs[name] = r.stdout
    assert outs["ours"] == outs["ref"]
    assert "flagged" in outs["ref"] and "run_new_md" in outs["ref"]
    files = sorted(os.listdir(tmp_path / "ref"))
    assert files == sorted(os.listdir(tmp_path / "ours")) and len(files) > n_qp
    match, mismatch, errors = filecmp.cmpfiles(tmp_path / "ref", tmp_path / "ours", files, shallow=False)
    assert not mismatch and not errors
    assert any(os.path.getsize(tmp_path / "ref" / f) > 0 for f in files if f.endswith("similar_hist"))


def test_in_process_pattern_uploads_only_the_new_samples(tmp_path):
    """FE_problem.h's pattern (append a sample to every point, fit every point, compare the flagged subset) keeps the
    histories on the GPU: one store build, then 48 bytes per point and NEW timestep — same bytes out as the reference,
    and as the flat-upload path (SCEMA_B200_STORE=0)."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref not built")
    ours_exe = build_ours(tmp_path)
    n_qp, n_steps = 300, 21
    args = [str(n_qp), str(n_steps), "10", "4e-7", "9"]
    outs, errs = {}, {}
    for name, exe, env in (("ref", REF_BIN, {}), ("store", ours_exe, {"DROPIN_STATS": "1"}),
                           ("flat", ours_exe, {"DROPIN_STATS": "1", "SCEMA_B200_STORE": "0"})):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, str(d)] + args, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        outs[name], errs[name] = r.stdout, r.stderr
    assert outs["store"] == outs["ref"] == outs["flat"]
    files = sorted(os.listdir(tmp_path / "ref"))
    for other in ("store", "flat"):
        match, mismatch, errors = filecmp.cmpfiles(tmp_path / "ref", tmp_path / other, files, shallow=False)
        assert not mismatch and not errors, other
    stats = dict(kv.split("=") for kv in errs["store"].split("store:")[1].split())
    # comparisons at t = 3, 5, 10, 15, 20, 21: one build (3 samples), then exactly the samples added since: 21 in total
    assert int(stats["rebuilds"]) == 1 and int(stats["appended_steps"]) == n_steps and int(stats["flat_uploads"]) == 0, stats
    assert int(stats["h2d_bytes"]) == 48 * n_qp * n_steps, stats
    flat = dict(kv.split("=") for kv in errs["flat"].split("store:")[1].split())
    assert int(flat["appended_steps"]) == 0 and int(flat["rebuilds"]) == 0


@pytest.mark.parametrize("legacy,env", [("1", {"SCEMA_B200_ALL_SIMILAR": "1"}), ("2", {"SCEMA_B200_NEAREST": "1"})])
def test_legacy_outputs_on_request(tmp_path, legacy, env):
    """The reference's "theory-checking" outputs: all_similar_histories_to_file (every comparison of every history,
    FE_problem.h:1237-1238) with SCEMA_B200_ALL_SIMILAR=1, and the legacy global nearest neighbour
    (get_most_similar_history_ID / _diff, lowest ID on ties) with either knob — byte-identical to the reference header."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref not built")
    ours_exe = build_ours(tmp_path)
    outs = {}
    for name, exe, extra in (("ref", REF_BIN, {}), ("ours", ours_exe, env)):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, str(d), "120", "12", "10", "1e-6", "4"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, DROPIN_LEGACY=legacy, **extra))
        assert r.returncode == 0, r.stderr
        outs[name] = r.stdout
    assert outs["ours"] == outs["ref"] and outs["ref"].count("nearest of") >= 40
    files = sorted(os.listdir(tmp_path / "ref"))
    assert files == sorted(os.listdir(tmp_path / "ours"))
    match, mismatch, errors = filecmp.cmpfiles(tmp_path / "ref", tmp_path / "ours", files, shallow=False)
    assert not mismatch and not errors
    if legacy == "1":
        full = [f for f in files if f.endswith("all_similar_hist")]
        assert full and all(os.path.getsize(tmp_path / "ref" / f) > 1000 for f in full)


def test_error_behaviour_matches(tmp_path):
    """< 3 samples: both print the reference's message and exit(1) (strain2spline.h:142-148)."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref not built")
    ours_exe = build_ours(tmp_path)
    for n_steps in ("2", "0"):
        res = []
        for exe in (REF_BIN, ours_exe):
            d = tmp_path / (os.path.basename(exe) + n_steps)
            d.mkdir()
            r = subprocess.run([exe, str(d), "16", n_steps, "10", "1e-6", "1"], capture_output=True, text=True, timeout=120)
            res.append((r.returncode, r.stdout, r.stderr))
        assert res[0] == res[1] and res[0][0] == 1 and "splinify" in res[0][2]
