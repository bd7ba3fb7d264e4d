"""Developer smoke on a real B200: FP64 peaks, K1/K2 parity vs the oracle, first timings."""
import json, os, subprocess, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scema_b200
from scema_b200 import synth, HistCluster, PAIRS_DMMA, PAIRS_FMA, PAIRS_EXACT
from oracle.pyoracle import Oracle

out = {}
def section(name):
    def deco(f):
        t = time.time()
        try:
            out[name] = f()
        except Exception as e:
            out[name] = {"error": repr(e), "tb": traceback.format_exc()}
        print(f"[{name}] {time.time()-t:.1f}s ->", json.dumps(out[name])[:2000], flush=True)
    return deco

o = Oracle()
hc = HistCluster(0)
thr, P = 1e-6, 10

@section("fp64_peak_tool")
def _():
    exe = os.path.join(os.path.dirname(__file__), "fp64_peak")
    return json.loads(subprocess.check_output([exe], timeout=120).decode())

@section("fp64_peak_lib")
def _():
    return hc.fp64_peak()

@section("k1_parity_ragged")
def _():
    off = synth.offsets(3, 3001, 16, 3, 200)
    st = synth.histories(3, 3001, 16, 5e-3, synth.default_pert(thr, P), off)
    want = o.splinify_batch(st, off, P)
    hc.set_histories(st, off)
    hc.resample(P)
    got = hc.get_spline()
    bad = int(np.count_nonzero(got.view(np.uint64) != want.view(np.uint64)))
    return {"n": 3001, "mismatch_doubles": bad, "max_abs": float(np.abs(got - want).max()), "t": hc.timings()}

@section("k2_parity")
def _():
    res = {}
    rows = synth.rows(2, 3000, 16, P, 5e-3, synth.default_pert(thr, P))
    # plant near-threshold pairs
    rng = np.random.default_rng(0)
    for k in range(200):
        a = rng.integers(0, 3000); b = (a + 1500) % 3000
        u = rng.standard_normal(60); u /= np.linalg.norm(u)
        rows[b] = rows[a] + u * thr * (1 + (k - 100) * 1e-15)
    wi, wj, wd, _ = o.all_pairs(rows, thr)
    for name, v in (("dmma", PAIRS_DMMA), ("fma", PAIRS_FMA), ("exact", PAIRS_EXACT)):
        hc.set_spline(rows)
        ne = hc.compare(thr, v)
        a, b, d = hc.get_edges()
        ok = len(a) == len(wi) and np.array_equal(a, wi) and np.array_equal(b, wj) and np.array_equal(d.view(np.uint64), wd.view(np.uint64))
        res[name] = {"edges": ne, "want": len(wi), "identical": bool(ok), "counters": hc.counters()}
    return res

def timing(n, variant, reps=3, Pp=10):
    rows = synth.rows(4, n, 16, Pp, 5e-3, synth.default_pert(thr, Pp))
    hc.set_spline(rows)
    best = None
    for r in range(reps):
        ne = hc.compare(thr, variant)
        t = hc.timings()
        if best is None or t["filter"] < best["filter"]:
            best = t
    pairs = n * (n - 1) / 2
    K = 6 * Pp
    return {"n": n, "edges": ne, "t_ms": best, "pairs_per_s_filter": pairs / (best["filter"] * 1e-3),
            "tflops_filter": pairs * 2 * K / (best["filter"] * 1e-3) / 1e12, "counters": hc.counters()}

for n in (16384, 65536, 200000):
    for name, v in (("dmma", PAIRS_DMMA), ("fma", PAIRS_FMA)):
        section(f"time_{name}_{n}")(lambda n=n, v=v: timing(n, v))
section("time_exact_16384")(lambda: timing(16384, PAIRS_EXACT))
section("time_exact_65536")(lambda: timing(65536, PAIRS_EXACT, reps=2))
section("time_dmma_p50_32768")(lambda: timing(32768, PAIRS_DMMA, Pp=50))
section("time_fma_p50_32768")(lambda: timing(32768, PAIRS_FMA, Pp=50))

@section("k1_timing_200k")
def _():
    n = 200000
    off = synth.offsets(3, n, 16, 6, 200)
    import ctypes
    st = synth.histories(3, n, 16, 5e-3, synth.default_pert(thr, P), off)
    hc.set_histories(st, off)
    ts = []
    for r in range(3):
        hc.resample(P)
        ts.append(hc.timings()["resample"])
    byts = int(off[-1]) * 48 + n * 480
    return {"n": n, "ms": ts, "GBps": byts / (min(ts) * 1e-3) / 1e9, "bytes": byts}

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_first.json", "w"), indent=1)
