"""Model-level check of the tcgen05 filter's guard band, on the CPU.

The operand preparation of scema_b200/csrc/pairs_tc.cu (k_tc_prep: power-of-two scale, fp16 slices with flushed
subnormals, fold columns carrying -h_i) is restated in numpy; the three sliced products are summed exactly (float64)
and then pushed DOWN by the largest error the design budgets for the tensor core's fp32 accumulation
((steps + 1) * 2^-18 * sum |terms|). Even then no pair the reference would call an edge may come out negative — for
both slice counts, on clustered rows, pairs planted a hair inside the threshold, rows of mixed magnitude, tiny and
huge thresholds. This pins the ARITHMETIC of the bound (constants c, e0, T', fold split); that the hardware stays
inside the budget is what tests/test_gpu_tc.py::test_tc_accumulators_match_sliced_fp64 measures on the GPU.
"""
import numpy as np
import pytest

P_, Q_ = 32768.0, 8.0


def h16z(v):
    """fp16 rounding (nearest even) with subnormal results flushed to zero, as the prep kernel does"""
    with np.errstate(over="ignore"):
        f = np.asarray(v, dtype=np.float64).astype(np.float16).astype(np.float64)
    return np.where(np.abs(f) < 2.0 ** -14, 0.0, f)


def centre_of(rows):
    """k_tc_centre_partial / _final: column means over a strided sample of <= 4096 rows, non-finite entries skipped
    (any vector would do for correctness: d(a, b) = d(a - m, b - m); the mean makes the centred norms small)"""
    n = len(rows)
    ns = min(n, 4096)
    sample = rows[(np.arange(ns) * (n // ns))]
    ok = np.isfinite(sample)
    cnt = ok.sum(0)
    with np.errstate(invalid="ignore", over="ignore"):
        m = np.where(cnt > 0, np.where(ok, sample, 0.0).sum(0) / np.maximum(cnt, 1), 0.0)
    return np.where(np.isfinite(m), m, 0.0)


def prepare(rows, thr, slices, centred=True):
    n, K = rows.shape
    assert K <= 60
    if centred:
        with np.errstate(invalid="ignore", over="ignore"):
            rows = rows - centre_of(rows)[None, :]      # fl(a - m), one rounding per element
    with np.errstate(over="ignore", invalid="ignore"):
        nrm_raw = (rows ** 2).sum(1)
    finite = np.isfinite(nrm_raw)
    M = np.abs(rows[finite]).max() if finite.any() else 0.0
    s = 2.0 ** (11 - int(np.floor(np.log2(M)))) if M > 0 else 1.0
    a = np.zeros((n, 60))
    a[:, :K] = np.where(finite[:, None], rows * s, 0.0)
    hi = h16z(a)
    lo = h16z(a - hi)
    cg = 2.0 ** -9 if slices == 1 else 2.0 ** -13
    eps = 2.0 ** -53
    T0 = thr * thr * (1 + (2 * K + 16) * eps) * (1 + 4 * eps) + (2 * K + 4) * 4.9406564584124654e-324
    T0s = (T0 * s) * s
    nrm = (a ** 2).sum(1)
    h = 0.5 * (nrm * (1 - cg) - T0s * (0.5 + 2 * cg) - K * 2.0 ** -7)
    force = (~finite) | ~(-h <= 32768.0 * P_)
    with np.errstate(invalid="ignore", over="ignore"):
        x0 = h16z(-h / P_)
        x1 = h16z((-h - P_ * x0) / Q_)
        x2 = h16z((-h - P_ * x0 - Q_ * x1) / Q_)
    x0 = np.where(force, 65504.0, x0)
    x1 = np.where(force, 0.0, x1)
    x2 = np.where(force, 0.0, x2)
    z = np.zeros(n)
    one = np.ones(n)
    a_hi = np.concatenate([hi, np.stack([x0, x1, P_ * one, Q_ * one], 1)], 1)
    a_lo = np.concatenate([lo, np.stack([z, x2, z, z], 1)], 1)
    b_hi = np.concatenate([hi, np.stack([P_ * one, Q_ * one, x0, x1], 1)], 1)
    b_lo = np.concatenate([lo, np.stack([z, z, z, x2], 1)], 1)
    return a_hi, a_lo, b_hi, b_lo


def worst_case_acc(rows, thr, slices, centred=True):
    a_hi, a_lo, b_hi, b_lo = prepare(rows, thr, slices, centred)
    acc = a_hi @ b_hi.T
    absum = np.abs(a_hi) @ np.abs(b_hi).T
    steps = 4
    if slices == 2:
        acc = acc + a_lo @ b_hi.T + a_hi @ b_lo.T
        absum = absum + np.abs(a_lo) @ np.abs(b_hi).T + np.abs(a_hi) @ np.abs(b_lo).T
        steps = 12
    return acc - (steps + 1) * 2.0 ** -18 * absum


def reference_edges(rows, thr):
    """compare_L2_norm in the reference's order (sequential k, separate multiply and add), strict threshold"""
    n, K = rows.shape
    s = np.zeros((n, n))
    for k in range(K):
        d = rows[:, None, k] - rows[None, :, k]
        s = s + d * d
    with np.errstate(invalid="ignore"):
        return np.sqrt(s) < thr


def cases():
    rng = np.random.default_rng(7)
    n = 400
    t = np.linspace(0, 1, 10)
    centres = rng.uniform(-5e-3, 5e-3, size=(n // 8, 6))
    base = (centres[:, None, :] * t[None, :, None]).reshape(n // 8, 60)
    clustered = np.repeat(base, 8, axis=0) + 3e-7 * rng.standard_normal((n, 60)) / np.sqrt(60)
    yield "clustered", clustered, 1e-6
    planted = clustered.copy()
    for q in range(0, n - 1, 2):  # partner a hair inside / outside the threshold
        u = rng.standard_normal(60)
        u /= np.linalg.norm(u)
        planted[q + 1] = planted[q] + u * 1e-6 * (1 + (q % 7 - 3) * 1e-13)
    yield "planted", planted, 1e-6
    mixed = rng.standard_normal((n, 60)) * 10.0 ** rng.integers(-7, 0, size=(n, 1))
    mixed[1::2] = mixed[::2] * (1 + 1e-9 * rng.standard_normal((n // 2, 60)))
    yield "mixed_magnitudes", mixed, 1e-7
    yield "tiny_threshold", clustered * 1e-6, 1e-12
    yield "large_threshold", clustered, 3e-3
    yield "k18", rng.standard_normal((n, 18)) * 1e-2 + 1.0, 6e-2
    # production shape: every history on nearly the same stretch path, groups ~1e-2 of the norm apart
    dq = 5e-3 * rng.uniform(-1, 1, size=n // 8)
    sig = 1e-2 * rng.uniform(-1, 1, size=(n // 8, 3))
    zz = 2e-2 * t[None, :] * (1 + dq[:, None])
    comp = np.stack([-0.3 * zz, -0.3 * zz, zz] + [zz * sig[:, c:c + 1] for c in range(3)], axis=2)   # [group, point, 6]
    smooth = np.repeat(comp.reshape(n // 8, 60), 8, axis=0) + 3e-7 * rng.standard_normal((n, 60)) / np.sqrt(60)
    yield "smooth", smooth, 1e-6
    yield "smooth_offset", smooth + 0.3, 1e-6     # a large common offset on top: only the centred copies can filter
    zero = clustered.copy()
    zero[:40] = 0.0
    zero[40:60] = 1e-30 * rng.standard_normal((20, 60))
    yield "zero_rows", zero, 1e-6


@pytest.mark.parametrize("slices", [1, 2])
@pytest.mark.parametrize("name,rows,thr", list(cases()), ids=[c[0] for c in cases()])
def test_no_edge_is_rejected_under_the_worst_budgeted_error(name, rows, thr, slices):
    edge = reference_edges(rows, thr)
    acc = worst_case_acc(rows, thr, slices)
    off_diag = ~np.eye(len(rows), dtype=bool)
    bad = edge & off_diag & (acc < 0)
    assert not bad.any(), (name, slices, int(bad.sum()), np.argwhere(bad)[:3])
    assert (edge & off_diag).sum() > 0
    # and the filter does reject most of what is far away (it is a filter, after all), except where the threshold
    # is of the order of the spread of the data
    if name in ("clustered", "planted", "tiny_threshold", "smooth", "smooth_offset"):
        assert (acc < 0).mean() > 0.9
    if name in ("smooth", "smooth_offset") and slices == 1:
        # ... which on production-shaped rows is owed to the centring: the band of the raw rows keeps nearly everything
        raw = worst_case_acc(rows, thr, slices, centred=False)
        assert not (edge & off_diag & (raw < 0)).any()
        assert (raw < 0).mean() < 0.2


def test_fold_columns_carry_minus_h():
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((200, 60)) * 10.0 ** rng.integers(-5, 0, size=(200, 1))
    for slices in (1, 2):
        a_hi, a_lo, b_hi, b_lo = prepare(rows, 1e-6, slices)
        fold_a = P_ * a_hi[:, 60] + Q_ * a_hi[:, 61] + (Q_ * a_lo[:, 61] if slices == 2 else 0.0)
        fold_b = P_ * b_hi[:, 62] + Q_ * b_hi[:, 63] + (Q_ * b_lo[:, 63] if slices == 2 else 0.0)
        assert np.array_equal(fold_a, fold_b)
        for arr in (a_hi, a_lo):
            nz = np.abs(arr[arr != 0])
            assert nz.min() >= 2.0 ** -14 and nz.max() <= 65504


def band_plan(rows, thr, rows_per_tile):
    """numpy restatement of k_tc_band_plan (pairs_tc.cu): rows sorted by squared norm; for every row tile the first
    256-row column tile whose smallest norm is out of reach of the row tile's largest one."""
    n, K = rows.shape
    nrm2 = (rows ** 2).sum(1)
    perm = np.argsort(nrm2, kind="stable")
    s2 = nrm2[perm]
    n_pad = (n + 255) // 256 * 256
    n_col = n_pad // 256
    rpc = 256 // rows_per_tile
    keps = (K + 8.0) * 2.0 ** -52
    jend = []
    for I in range(n_pad // rows_per_tile):
        last = min((I + 1) * rows_per_tile, n) - 1
        hi = np.sqrt(s2[last])
        je = n_col
        for J in range(I // rpc + 1, n_col):
            l = np.sqrt(s2[J * 256])
            if l - hi > thr * (1 + 4 * keps) + 4 * keps * l + 1e-150:
                je = J
                break
        jend.append(je)
    return perm, jend


@pytest.mark.parametrize("rows_per_tile", [128, 256])
def test_norm_band_never_cuts_an_edge(rows_per_tile):
    """Every pair the reference calls an edge lies inside the band of tiles the norm-band schedule walks."""
    rng = np.random.default_rng(5)
    n = 1500
    for name, rows, thr in (
        ("spread", rng.standard_normal((n, 60)) * 10.0 ** rng.uniform(-4, -2, size=(n, 1)), 1e-6),
        ("near", 1e-3 * (1 + 1e-4 * rng.standard_normal((n, 60))), 3e-6),       # many edges, norms within a few thr
        ("tiny", 1e-160 * rng.standard_normal((n, 60)), 1e-159),               # squares underflow in the norms
        ("equal", np.tile(rng.standard_normal((1, 60)), (n, 1)) * 1e-2, 1e-6),   # all distances 0
    ):
        if name == "spread":
            rows[1::2] = rows[::2] + thr * 0.9 * rng.standard_normal((n // 2, 60)) / np.sqrt(60)
        edge = reference_edges(rows, thr)
        perm, jend = band_plan(rows, thr, rows_per_tile)
        pos = np.empty(n, dtype=np.int64)
        pos[perm] = np.arange(n)
        ii, jj = np.nonzero(np.triu(edge, 1))
        assert len(ii) > 0, name
        a = np.minimum(pos[ii], pos[jj])
        b = np.maximum(pos[ii], pos[jj])
        row_tile = a // rows_per_tile
        col_tile = b // 256
        limit = np.asarray(jend)[row_tile]
        assert np.all(col_tile < limit), (name, int(np.sum(col_tile >= limit)))
        if name == "spread":
            full = sum(len(jend) * 0 + (len(jend) // (256 // rows_per_tile)) for _ in [0])
            assert sum(je - I // (256 // rows_per_tile) for I, je in enumerate(jend)) < 0.6 * len(jend) * full  # it does prune


def prepare_wide(rows, thr):
    """k_tc_prep for rows of several 64-column chunks (hi slices only): scale one binade lower for K > 64, two for
    K > 256; the fold columns sit in the last four columns of the last chunk."""
    n, K = rows.shape
    nc = (K + 4 + 63) // 64
    assert 1 < nc <= 10
    hb = 0 if K <= 64 else (1 if K <= 256 else 2)
    rows = rows - centre_of(rows)[None, :]
    M = np.abs(rows).max()
    s = 2.0 ** (11 - hb - int(np.floor(np.log2(M))))
    a = np.zeros((n, nc * 64))
    a[:, :K] = rows * s
    assert np.abs(a).max() < 2.0 ** (12 - hb)
    hi = h16z(a)
    cg = 2.0 ** -9
    eps = 2.0 ** -53
    T0 = thr * thr * (1 + (2 * K + 16) * eps) * (1 + 4 * eps) + (2 * K + 4) * 4.9406564584124654e-324
    nrm = (a ** 2).sum(1)
    assert nrm.max() < 2.0 ** 30
    h = 0.5 * (nrm * (1 - cg) - (T0 * s) * s * (0.5 + 2 * cg) - K * 2.0 ** -7)
    force = ~(-h <= 32768.0 * P_)
    x0 = np.where(force, 65504.0, h16z(-h / P_))
    x1 = np.where(force, 0.0, h16z((-h - P_ * x0) / Q_))
    one = np.ones(n)
    a_hi, b_hi = hi.copy(), hi.copy()
    a_hi[:, -4:] = np.stack([x0, x1, P_ * one, Q_ * one], 1)
    b_hi[:, -4:] = np.stack([P_ * one, Q_ * one, x0, x1], 1)
    return a_hi, b_hi, nc


@pytest.mark.parametrize("K", [66, 300, 636])
def test_wide_rows_no_edge_is_rejected_under_the_worst_budgeted_error(K):
    rng = np.random.default_rng(K)
    n = 300
    t = np.linspace(0, 1, K // 6)
    centres = rng.uniform(-5e-3, 5e-3, size=(n // 6, 6))
    base = (centres[:, None, :] * t[None, :, None]).reshape(n // 6, K)
    rows = np.repeat(base, 6, axis=0) + 3e-7 * rng.standard_normal((n, K)) / np.sqrt(K)
    for q in range(0, n - 1, 3):  # partners a hair inside the threshold
        u = rng.standard_normal(K)
        u /= np.linalg.norm(u)
        rows[q + 1] = rows[q] + u * 1e-6 * (1 - 1e-13)
    thr = 1e-6
    edge = reference_edges(rows, thr)
    a_hi, b_hi, nc = prepare_wide(rows, thr)
    assert K > 60 and a_hi.shape[1] - 4 >= K                     # the data never reaches into the fold columns
    acc = a_hi @ b_hi.T - (4 * nc + 1) * 2.0 ** -18 * (np.abs(a_hi) @ np.abs(b_hi).T)
    off_diag = ~np.eye(n, dtype=bool)
    assert (edge & off_diag).sum() > n // 3
    assert not (edge & off_diag & (acc < 0)).any()
    assert (acc < 0).mean() > 0.9
