"""k_resample_pair (scema_b200/csrc/resample_pair.cuh) run on the CPU: tests/helpers/k1_emul.cpp compiles the kernel
SOURCE for the host, runs it thread by thread (128 threads + a barrier per CTA, shared memory as a byte array, cp.async
landing at the earliest / the latest legal moment / a random mix) and the result must equal the CPU oracle bit for bit
(reference: strain2spline.h:140-180 over spline.h:187-396). Checks the indexing of the prefetch rings, the peeled first
and last steps, the pairing of groups, partial groups, the chunk hand-out and the slow-path hand-over without a GPU;
the GPU parity tests (test_gpu_parity.py::test_k1_*) remain the proof for the compiled kernel."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"eager": 0, "lazy": 1, "mixed": 2}


@pytest.fixture(scope="module", params=[0, 16], ids=["depth-default", "depth-16"])
def emul(tmp_path_factory, request):
    """The emulation library, built with the kernel's own ring depth (8 slots per chain) and with 16 (PR_DEPTH_SLOTS)."""
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA instruction")
    so = str(tmp_path_factory.mktemp("k1emul") / "libk1emul.so")
    depth = ["-DPR_DEPTH_SLOTS=%d" % request.param] if request.param else []
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-Wno-unknown-pragmas"] + depth + ["-o", so, os.path.join(HERE, "helpers", "k1_emul.cpp")])
    lib = ctypes.CDLL(so)
    dp, u64p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)
    lib.k1_emul_ragged.argtypes = [dp, u64p, ctypes.c_uint64, ctypes.c_uint32, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, u64p]
    lib.k1_emul_store.argtypes = [dp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, dp, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, u64p]
    return lib


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def run_ragged(lib, steps, off, P, mode, n_ctas=2, global_table=False, flags=0):
    steps = np.ascontiguousarray(steps, dtype=np.float64)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    out = np.full((n, 6 * P), np.nan)
    stats = np.zeros(3, dtype=np.uint64)
    rc = lib.k1_emul_ragged(_ptr(steps, ctypes.c_double), _ptr(off, ctypes.c_uint64), n, P, _ptr(out, ctypes.c_double), MODES[mode], n_ctas,
                            int(global_table), flags, _ptr(stats, ctypes.c_uint64))
    assert rc == 0
    assert stats[2] == 0, "the kernel touched memory outside its buffers / its own ring"
    return out, stats


def same_bits(a, b):
    """Same bits; any NaN equals any NaN (the payload / sign of a generated NaN differs between x86 and the GPU, as in
    tests/test_gpu_parity.py)."""
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("mode", list(MODES))
def test_every_short_length_and_partial_groups(emul, oracle, mode):
    """L = 3 .. 40 (all residues of the two-step loops, the ring prologue shorter and longer than the history) with 1 .. 23
    histories per length: full pairs, a lone group, partial groups on either side of a pair, several chunks."""
    rng = np.random.default_rng(11)
    lens = []
    for L in range(3, 41):
        lens += [L] * int(rng.integers(1, 24))
    lens += [17] * 170  # 34 groups of one length: three chunks, the last one with two groups
    lens = np.array(lens, dtype=np.uint64)
    rng.shuffle(lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    steps = rng.standard_normal((int(off[-1]), 6)) * 1e-3
    for P in (1, 2, 10, 50):
        got, stats = run_ragged(emul, steps, off, P, mode, n_ctas=3)
        with np.errstate(all="ignore"):
            want = oracle.splinify_batch(steps, off, P)  # P = 1: 0/0 abscissa, NaN samples (strain2spline.h:171)
        assert same_bits(got, want), P
        assert stats[0] > 0
        if mode == "lazy":
            assert stats[1] == stats[0]  # every copy landed only when a wait forced it


# memory-behaviour flags of the kernel (resample_pair.cuh): y prefetch whole / none / windows, evict_first on the y copies —
# the result may not depend on any of them
FLAGS = [0, 1, 2, 2 | 4]


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("mode", list(MODES))
def test_longer_lengths_both_table_paths(emul, oracle, mode, flags):
    """Lengths across the class boundaries (64 | 65, 256 | 257: table in shared memory / read from global memory), odd and
    even, and the same batch with the shared-memory table switched off."""
    rng = np.random.default_rng(12)
    lens = np.array([63, 64, 64, 65, 66, 100, 101, 199, 200, 255, 256, 256, 257, 258, 300, 64, 64, 64, 64, 64, 64, 64, 200, 200,
                     200, 200, 200, 200, 3, 4, 5], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    steps = rng.standard_normal((int(off[-1]), 6)) * 1e-2
    want = oracle.splinify_batch(steps, off, 10)
    got, stats = run_ragged(emul, steps, off, 10, mode, flags=flags)
    assert same_bits(got, want)
    got, _ = run_ragged(emul, steps, off, 10, mode, global_table=True, flags=flags)
    assert same_bits(got, want)


@pytest.mark.parametrize("flags", [0, 2])
@pytest.mark.parametrize("mode", ["eager", "lazy"])
def test_special_magnitudes_take_the_slow_path(emul, oracle, mode, flags):
    """The data of test_gpu_parity.py::test_k1_special_magnitudes: zeros and -0 stay on the fast path (sign of a zero
    quotient), subnormal / huge / inf / NaN numerators hand the chain to the IEEE-division path, chain A and chain B of a
    lane independently."""
    n, L = 64, 37
    off = (np.arange(n + 1, dtype=np.uint64) * L)
    rng = np.random.default_rng(5)
    steps = rng.standard_normal((n * L, 6)) * 1e-3
    scale = np.array([1e-300, 1e-310, 5e-324, 1e300, 1e-280, 1e290, 2.0 ** -895, 2.0 ** 897])
    for q in range(8):
        steps[q * L:(q + 1) * L] *= scale[q]
    steps[8 * L + 3, 1] = np.inf
    steps[9 * L + 5, 2] = np.nan
    steps[10 * L:(11 * L), 0] = 0.0
    steps[11 * L:(12 * L), 3] = -0.0
    steps[12 * L + 7, 4] = 5e-324
    steps[13 * L:(14 * L), 5] = 1.7976931348623157e308
    steps[20 * L:(21 * L)] = 0.0           # a whole history of zeros
    steps[21 * L:(22 * L)] = -0.0
    steps[22 * L:(23 * L), 2] *= -1.0
    for P in (10, 50):  # P > L puts several samples into one interval
        got, _ = run_ragged(emul, steps, off, P, mode, flags=flags)
        with np.errstate(all="ignore"):
            want = oracle.splinify_batch(steps, off, P)
        assert same_bits(got, want), P


@pytest.mark.parametrize("flags", [0, 2])
@pytest.mark.parametrize("mode", list(MODES))
def test_history_store_layout(emul, oracle, mode, flags):
    """Time-major store [step][n][6] (implicit chunks, groups in index order, n not a multiple of 5 or 10)."""
    rng = np.random.default_rng(13)
    for n, L in ((1, 3), (7, 9), (83, 20), (161, 12), (10, 70)):
        tm = rng.standard_normal((L, n, 6)) * 1e-4
        out = np.full((n, 60), np.nan)
        stats = np.zeros(3, dtype=np.uint64)
        rc = emul.k1_emul_store(_ptr(tm, ctypes.c_double), n, L, 10, _ptr(out, ctypes.c_double), MODES[mode], 2, flags, _ptr(stats, ctypes.c_uint64))
        assert rc == 0 and stats[2] == 0
        ragged = np.ascontiguousarray(tm.transpose(1, 0, 2)).reshape(n * L, 6)
        off = (np.arange(n + 1, dtype=np.uint64) * L)
        assert same_bits(out, oracle.splinify_batch(ragged, off, 10)), (n, L)


def test_golden_kats(emul):
    """The committed known-answer vectors (tests/golden/kat_spline.json, generated by the unmodified reference header)."""
    import json
    for kat in json.load(open(os.path.join(HERE, "golden", "kat_spline.json"))):
        steps = np.array([float.fromhex(v) for v in kat["steps"]]).reshape(kat["L"], 6)
        off = np.array([0, len(steps)], dtype=np.uint64)
        got, _ = run_ragged(emul, steps, off, kat["P"], "lazy")
        want = np.array([float.fromhex(v) for v in kat["spline"]])
        assert same_bits(got[0], want), kat["name"]
