"""Diagnostic for the tcgen05 filter (SCEMA_PAIRS_TC): dumps every accumulator of a small problem and
compares it with (1) the same sliced contraction redone in FP64 from the fp16 operand copies
(layout / descriptor / accumulation check) and (2) s^2 a.b - h_i - h_j from the FP64 rows, then
checks the edge list against the DMMA variant and times a larger problem.
Run on a GPU box:  SCEMA_TC_CG=2 python tools/tc_probe.py [n_small] [n_big]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scema_b200  # noqa: E402
from scema_b200 import binding, synth  # noqa: E402


def unswizzle(buf, n_pad):
    """operand bytes -> (hi, lo) float64 arrays [n_pad, 64]"""
    b = buf.reshape(n_pad // 128, 2, 128, 8, 16)  # block, slice, row, stored chunk, bytes
    out = np.empty_like(b)
    r = np.arange(128)
    for c in range(8):
        out[:, :, r, c, :] = b[:, :, r, c ^ (r & 7), :]
    h = out.reshape(n_pad // 128, 2, 128, 128).view(np.float16).astype(np.float64)  # [blk, slice, row, 64]
    hi = h[:, 0].reshape(n_pad, 64)
    lo = h[:, 1].reshape(n_pad, 64)
    return hi, lo


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    thr, P = 1e-6, 10
    cg = os.environ.get("SCEMA_TC_CG", "2")
    print(f"== tc_probe CG={cg} slices={os.environ.get('SCEMA_TC_SLICES', 'auto')} n={n}", flush=True)
    rows = synth.rows(11, n, 16, P, 5e-3, synth.default_pert(thr, P))
    hc = scema_b200.HistCluster(0)
    hc.set_spline(rows)
    slices = int(os.environ.get("SCEMA_TC_SLICES", "2"))
    acc, ha, hb = hc.tc_debug(thr, n, slices)
    n_pad = acc.shape[0]
    ahi, alo = unswizzle(ha, n_pad)
    bhi, blo = unswizzle(hb, n_pad)
    want = ahi @ bhi.T
    absum = np.abs(ahi) @ np.abs(bhi).T
    if slices == 2:
        want = want + alo @ bhi.T + ahi @ blo.T
        absum = absum + np.abs(alo) @ np.abs(bhi).T + np.abs(ahi) @ np.abs(blo).T
    written = ~np.isnan(acc)
    # which tiles must have been written: column tile J >= row tile I (256-row units)
    ti = np.arange(n_pad)[:, None] // 256
    tj = np.arange(n_pad)[None, :] // 256
    must = tj >= ti
    print("written where required:", bool(np.all(written[must])), " written elsewhere:", int(np.sum(written & ~must)))
    real = must & written & (np.arange(n_pad)[:, None] < n) & (np.arange(n_pad)[None, :] < n)
    err = np.abs(acc.astype(np.float64) - want)
    rel = np.where(real, err / np.maximum(absum, 1e-300), 0.0)
    steps = 12 if slices == 2 else 4
    print("max |acc - fp64(sliced)| / sum|terms| = %.3e (2^%.1f); the guard band budgets %d*2^-18 = %.3e" %
          (rel.max(), np.log2(max(rel.max(), 1e-300)), steps + 1, (steps + 1) * 2.0 ** -18))
    i, j = np.unravel_index(np.argmax(rel), rel.shape)
    print("  worst at", (i, j), "acc", acc[i, j], "want", want[i, j], "absum", absum[i, j])
    for (i, j) in [(0, 1), (0, 17), (5, 300 % n), (n - 2, n - 1)]:
        if i < n and j < n and must[i, j]:
            print("  sample", (i, j), "acc", acc[i, j], "want", want[i, j])
    # sign decisions vs the truth
    d2 = None
    if n <= 4096:
        d2 = np.empty((n, n))
        for r0 in range(0, n, 128):
            d2[r0:r0 + 128] = ((rows[r0:r0 + 128, None, :] - rows[None, :, :]) ** 2).sum(-1)
    if d2 is not None:
        surv = (acc[:n, :n] >= 0) | np.isnan(acc[:n, :n])
        edge = (np.sqrt(d2) < thr) & (np.arange(n)[:, None] < np.arange(n)[None, :])
        missed = edge & ~surv & must[:n, :n]
        print("true edges:", int(edge.sum()), "survivors (upper):", int((surv & (np.arange(n)[:, None] < np.arange(n)[None, :])).sum()),
              "missed edges:", int(missed.sum()))
    # full pipeline parity against the DMMA variant
    for m in (n, 5000):
        r = synth.rows(12, m, 16, P, 5e-3, synth.default_pert(thr, P))
        hc.set_spline(r)
        ne_t = hc.compare(thr, binding.PAIRS_TC)
        et = hc.get_edges()
        ct = hc.counters()
        ne_d = hc.compare(thr, binding.PAIRS_DMMA)
        ed = hc.get_edges()
        same = ne_t == ne_d and all(np.array_equal(x.view(np.uint64) if x.dtype == np.float64 else x,
                                                   y.view(np.uint64) if y.dtype == np.float64 else y) for x, y in zip(et, ed))
        print(f"parity n={m}: tc edges {ne_t} dmma edges {ne_d} identical={same} survivors={ct['survivors']}", flush=True)
    if n_big:
        import torch
        d_rows = synth.device_rows(4, n_big, 16, P, 5e-3, synth.default_pert(thr, P))
        hc.set_spline(device_ptr=d_rows.data_ptr(), n=n_big, k=6 * P)
        for variant, name in ((binding.PAIRS_TC, "tc"), (binding.PAIRS_TC, "tc"), (binding.PAIRS_DMMA, "dmma")):
            t0 = time.time()
            ne = hc.compare(thr, variant)
            torch.cuda.synchronize()
            wall = time.time() - t0
            tm = hc.timings()
            c = hc.counters()
            pairs = n_big * (n_big - 1) / 2
            print(f"big n={n_big} {name}: edges {ne} survivors {c['survivors']} filter {tm['filter']:.2f} ms prep {tm['prep']:.2f} ms "
                  f"slices {c['tc_slices']} passes {c['passes']} exact {tm['exact']:.2f} ms wall {wall*1e3:.1f} ms -> {pairs / (tm['filter'] * 1e-3):.3e} pairs/s (filter)", flush=True)
    hc.close()
    print("== done", flush=True)


if __name__ == "__main__":
    main()
