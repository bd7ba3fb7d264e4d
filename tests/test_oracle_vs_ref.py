"""The oracle (oracle/hist_oracle.c) against the golden vectors and, where built, the reference.

The reference has no tests of its own; the fixtures in tests/golden/ were produced by
tests/golden/make_golden.py from the unmodified reference header. Bit-exact comparisons only.
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(lst):
    return np.array([float.fromhex(x) for x in lst], dtype=np.float64)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    # NaN payload/sign is not specified by the reference either: treat all NaNs as one value
    return a.shape == b.shape and bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def load_kats():
    return json.load(open(os.path.join(GOLD, "kat_spline.json")))


@pytest.mark.parametrize("kat", load_kats(), ids=lambda k: k["name"])
def test_oracle_spline_golden(oracle, kat):
    steps = unhex(kat["steps"]).reshape(kat["L"], 6)
    got = oracle.splinify(steps, kat["P"])
    assert same_bits(got, unhex(kat["spline"]))


def test_survey_kat_values():
    """Hex constants quoted in SURVEY.md §8c (KAT-1, KAT-4) are what the fixture holds."""
    kats = {k["name"]: k for k in load_kats()}
    k1 = kats["kat1_L5_P4"]["spline"]
    assert k1[6] == "0x1.cc7c5a8ba1e67p-10" and k1[20] == "0x1.0624dd2f1a9fep-11" and k1[1] == "-0x0.0p+0"
    k4 = kats["kat4_L3_P10"]["spline"]
    assert k4[6 * 1] == "0x1.04b4a3c8a7ff7p-12" and k4[6 * 9] == "0x1.89374bc6a7efap-10"


def test_oracle_pairs_golden(oracle):
    g = json.load(open(os.path.join(GOLD, "kat_pairs.json")))
    rows = unhex(g["rows"]).reshape(g["n"], g["K"])
    for e in g["l2"]:
        assert oracle.compare_l2(rows[e["a"]], rows[e["b"]]).hex() == e["d"]
    ei, ej, ed, pairs = oracle.all_pairs(rows, g["thr"])
    assert pairs == g["n"] * (g["n"] - 1) // 2
    assert ei.tolist() == g["edges"]["i"] and ej.tolist() == g["edges"]["j"]
    assert same_bits(ed, unhex(g["edges"]["d"]))
    # the planted pairs straddle the threshold: both outcomes occur among them
    planted = [(k, 48 + k) for k in range(12)]
    found = set(zip(ei.tolist(), ej.tolist()))
    hits = sum(p in found for p in planted)
    assert 0 < hits < 12


def test_oracle_min_steps(oracle):
    with pytest.raises(ValueError):
        oracle.splinify(np.zeros((2, 6)), 10)  # reference: exit(1), strain2spline.h:145-148


def test_oracle_files_golden(oracle, tmp_path):
    """Per-history file contents == the reference CLI's __results/ID_<id>.txt (as line sets; the
    line ORDER follows the directory enumeration of the generating run and is checked live in
    test_cli.py where oracle/_ref exists)."""
    from scema_b200 import synth
    g = json.load(open(os.path.join(GOLD, "pipeline_c1", "reference_outputs.json")))
    n, P, L, thr = g["n"], g["P"], g["L"], g["thr"]
    off = synth.offsets(g["seed"], n, g["cluster"], L, L)
    steps = synth.histories(g["seed"], n, g["cluster"], g["amp"], synth.default_pert(thr, P), off)
    # the CLI parsed repr()-printed text, which round-trips doubles exactly
    sp = oracle.splinify_batch(steps, off, P)
    ei, ej, ed, _ = oracle.all_pairs(sp, thr)
    ids = np.arange(n, dtype=np.uint32)
    assert oracle.write_similar_files(ids, ei, ej, ed, str(tmp_path / "ID_%u.txt")) == 0
    for i in range(n):
        got = open(tmp_path / f"ID_{i}.txt").read()
        assert sorted(got.splitlines()) == sorted(g["results"][str(i)].splitlines()), i
    assert 2 * len(ei) == sum(len(v.splitlines()) for v in g["results"].values())


def test_oracle_vs_reference_random(oracle, reference):
    rng = np.random.default_rng(7)
    for trial in range(400):
        L = int(rng.integers(3, 260))
        P = int(rng.integers(1, 62))
        st = rng.standard_normal((L, 6)) * 10.0 ** rng.integers(-9, 1, size=6)
        if trial % 7 == 0:
            st[:, 2] = 0
        if trial % 11 == 0:
            st[:, 3] = st[0, 3]
        if trial % 13 == 0:
            st[:, 1] = -0.0
        assert same_bits(oracle.splinify(st, P), reference.splinify(st, P))
    rows = rng.standard_normal((400, 300)) * 1e-3
    rows[200:] = rows[:200] + rng.standard_normal((200, 300)) * 5.5e-8
    a = oracle.all_pairs(rows, 1e-6)
    b = reference.all_pairs(rows, 1e-6)
    assert len(a[0]) > 10 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and same_bits(a[2], b[2])
    for v in (1e-6, 0.0, 1.23456789e-7, 12345678.0, 1e-320, float("inf"), 9.9999995e-7, 100000.0, 999999.5, 1e22):
        assert oracle.format_double(v) == reference.format_double(v)


def test_reference_pipeline_matches_oracle(oracle, reference, tmp_path):
    """compare_histories_with_all_ranks + most_similar_histories_to_file (the as-called path)."""
    from scema_b200 import synth
    n, P, thr = 200, 10, 1e-6
    off = synth.offsets(5, n, 8, 4, 40)
    steps = synth.histories(5, n, 8, 1e-3, synth.default_pert(thr, P), off)
    ids = (np.arange(n, dtype=np.uint32) * 3 + 7)
    (tmp_path / "r").mkdir()
    (tmp_path / "o").mkdir()
    sp = reference.pipeline(steps, off, ids, P, thr, str(tmp_path / "r" / "last.%u.similar_hist"), want_spline=True)
    mine = oracle.splinify_batch(steps, off, P)
    assert same_bits(sp, mine)
    ei, ej, ed, _ = oracle.all_pairs(mine, thr)
    oracle.write_similar_files(ids, ei, ej, ed, str(tmp_path / "o" / "last.%u.similar_hist"))
    for i in ids:
        assert open(tmp_path / "r" / f"last.{i}.similar_hist").read() == open(tmp_path / "o" / f"last.{i}.similar_hist").read()
