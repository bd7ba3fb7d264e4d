#!/usr/bin/env python
"""K1 (ragged spline fit / resample) alone: time and algorithmic HBM bandwidth (48 bytes per raw step read + 8 K bytes per
history written) on the bench workloads (ragged batch path) and on the production shape (device-resident history store:
every quadrature point equally long, FE_problem.h:1167-1191 re-fits ALL points every timestep).
  python tools/k1_probe.py [n_store] [L_store]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scema_b200
from scema_b200 import synth


def measure(fn, hc, reps):
    ts = []
    for _ in range(reps):
        fn()
        ts.append(hc.timings()["resample"])
    return min(ts[1:])


def scan(hc, peak, n_store, L_store):
    """--scan: every case once in memory, then the tuning grid of scema_k1_tune in this one process (kernel x warps per SM x
    memory flags) — one line per setting, the best per case at the end."""
    P = 10
    cases = []
    for name, n, lmin, lmax in (("c3", 200000, 6, 200), ("c4", 1000000, 8, 64)):
        off = synth.device_offsets(4, n, 16, lmin, lmax)
        d = synth.device_histories(4, n, 16, 5e-3, 1e-7, off)
        cases.append((name, off, d, int(off[-1]) * 48 + n * 6 * P * 8))
    grid = [("stream", 0, 0, 0)] + [("pair", 1, w, f) for f in (0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14) for w in (8, 10, 12, 14, 16, 18, 20, 24)]
    rows = []
    for name, off, d, b in cases:
        hc.set_histories(None, off, device_ptr=d.data_ptr())
        for kname, k, w, f in grid:
            hc.k1_tune(kernel=k, wps_ragged=w, wps_store=w, flags=f)
            ms = measure(lambda: hc.resample(P), hc, 4)
            rows.append({"case": name, "kernel": kname, "wps": w, "flags": f, "ms": ms, "frac_of_hbm_peak": b / ms / 1e6 / peak})
    hc.store_reset(n_store, None, L_store)
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.rand((n_store, 6), device="cuda", dtype=torch.float64, generator=g) * 1e-6
    for s_ in range(L_store):
        step = base * (s_ + 1) + 1e-9 * torch.rand((n_store, 6), device="cuda", dtype=torch.float64, generator=g)
        torch.cuda.synchronize()
        hc.store_append(device_ptr=step.data_ptr())
    b = n_store * L_store * 48 + n_store * 6 * P * 8
    for kname, k, w, f in [("stream", 0, 0, 0)] + [("pair", 1, w, f) for f in (0, 4, 8, 12) for w in (12, 16, 20, 24)]:
        hc.k1_tune(kernel=k, wps_ragged=w, wps_store=w, flags=f)
        ms = measure(lambda: hc.store_resample(P), hc, 3)
        rows.append({"case": "store%dx%d" % (n_store, L_store), "kernel": kname, "wps": w, "flags": f, "ms": ms,
                     "frac_of_hbm_peak": b / ms / 1e6 / peak})
    print(json.dumps({"hbm_peak_gbs": peak, "lib": os.environ.get("SCEMA_LIB"), "rows": rows}))
    for r in rows:
        print("%-14s %-6s wps %2d flags %d  %8.3f ms  %.3f of HBM peak" % (r["case"], r["kernel"], r["wps"], r["flags"], r["ms"],
                                                                           r["frac_of_hbm_peak"]), file=sys.stderr)
    for c in sorted({r["case"] for r in rows}):
        best = min((r for r in rows if r["case"] == c), key=lambda r: r["ms"])
        print("best %-14s %s wps %d flags %d: %.3f ms" % (c, best["kernel"], best["wps"], best["flags"], best["ms"]), file=sys.stderr)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--scan" in sys.argv:
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6536.0
        hc = scema_b200.HistCluster(0)
        scan(hc, peak, int(args[0]) if args else 200000, int(args[1]) if len(args) > 1 else 500)
        hc.close()
        return
    n_store = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    L_store = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6536.0
    hc = scema_b200.HistCluster(0)
    out = {"hbm_peak_gbs": peak, "env": {k: os.environ.get(k) for k in ("SCEMA_K1_KERNEL", "SCEMA_K1_WPS", "SCEMA_LIB")}, "cases": []}
    P = 10
    for name, n, lmin, lmax in (("c3 ragged 6..200", 200000, 6, 200), ("c4 ragged 8..64", 1000000, 8, 64)):
        off = synth.device_offsets(4, n, 16, lmin, lmax)
        d = synth.device_histories(4, n, 16, 5e-3, 1e-7, off)
        hc.set_histories(None, off, device_ptr=d.data_ptr())
        ts = []
        for _ in range(6):
            hc.resample(P)
            ts.append(hc.timings()["resample"])
        ms = min(ts[2:])
        b = int(off[-1]) * 48 + n * 6 * P * 8
        out["cases"].append({"case": name, "histories": n, "ms": ms, "algorithmic_gb": b / 1e9, "gbs": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / peak})
        del d
    # production shape through the history store
    hc.store_reset(n_store, None, L_store)
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.rand((n_store, 6), device="cuda", dtype=torch.float64, generator=g) * 1e-6
    for s in range(L_store):
        step = base * (s + 1) + 1e-9 * torch.rand((n_store, 6), device="cuda", dtype=torch.float64, generator=g)
        torch.cuda.synchronize()
        hc.store_append(device_ptr=step.data_ptr())
    ts = []
    for _ in range(5):
        hc.store_resample(P)
        ts.append(hc.timings()["resample"])
    ms = min(ts[1:])
    b = n_store * L_store * 48 + n_store * 6 * P * 8
    out["cases"].append({"case": "store %d x %d steps (production shape)" % (n_store, L_store), "histories": n_store, "ms": ms,
                         "algorithmic_gb": b / 1e9, "gbs": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / peak})
    print(json.dumps(out))
    for c in out["cases"]:
        print("%-45s %8.3f ms  %7.1f GB/s  %.3f of HBM peak" % (c["case"], c["ms"], c["gbs"], c["frac_of_hbm_peak"]), file=sys.stderr)
    hc.close()


if __name__ == "__main__":
    main()
