"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
It drives oracle/_ref/libscema_ref.so (reference headers/strain2spline.h + spline.h compiled by
oracle/Makefile), oracle/_ref/mpi_comparison_test, oracle/_ref/compare_all_histories and the
reference's coarsegrain_dependency_network.py, and records their outputs:

  kat_spline.json    Strain6D::splinify outputs (hex floats) for hand-written and seeded inputs
  kat_pairs.json     compare_L2_norm values + the thresholded edge list of a small seeded matrix
  pipeline_c1/       a config-1-like case (576 histories, P=10, thr=1e-6) as strain_<ID> text files'
                     digest, every __results/ID_<id>.txt produced by the reference CLI, the
                     compare_all_histories stdout digest, and the mapping.csv of the python script
  graph_cases.json   add_edge call sequences and the mapping the real script produced
The reference ships no golden vectors of its own, so these are the pin.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference, ref_binary, build  # noqa: E402
from scema_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_PY = "/root/reference/clustering/coarsegrain_dependency_network.py"


def hexlist(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def kat_inputs():
    cases = []
    t = np.arange(5.0)
    xx = np.array([0, 1, 4, 9, 16.0]) * 1e-3
    cases.append(("kat1_L5_P4", np.stack([xx, -0.3 * xx, np.array([1, -2, 3, -4, 5.0]) * 1e-4, 2e-4 * t, 0 * t,
                                          np.full(5, 7e-5)], 1), 4))
    t = np.arange(6.0)
    cases.append(("kat2_L6_P4", np.stack([3.1e-3 * t * t / 5.0, -1e-3 * t, 1e-4 * np.where(t % 2 == 1, -1.0, 1.0),
                                          2.5e-4 * t, 1e-6 * t * t * t, np.full(6, 7e-5)], 1), 4))
    cases.append(("kat4_L3_P10", np.stack([np.array([0, 1e-3, 1.5e-3])] + [np.zeros(3)] * 5, 1), 10))
    cases.append(("negzero_L4_P3", np.stack([np.full(4, -0.0), np.array([1e-9, -2e-9, 3e-9, -4e-9]),
                                             np.zeros(4), np.full(4, 1e-300), np.array([1e300, -1e300, 1e300, 0.0]),
                                             np.array([0.0, 5e-324, 0.0, -5e-324])], 1), 3))
    cases.append(("P1_nan_L5", np.stack([np.arange(5.0)] * 6, 1), 1))
    cases.append(("P2_L3", np.stack([np.array([1.0, 2.0, 4.0])] * 6, 1) * 1e-3, 2))
    rng = np.random.default_rng(20261017)
    for L, P in ((3, 7), (4, 10), (7, 10), (10, 10), (19, 10), (37, 10), (50, 50), (100, 10), (200, 10), (201, 61),
                 (1000, 10), (64, 33)):
        st = rng.standard_normal((L, 6)) * 10.0 ** rng.integers(-8, -1, size=6)
        cases.append((f"rand_L{L}_P{P}", st, P))
    return cases


def main():
    build(ref=True)
    ref = Reference()
    # ---- spline KATs
    kat = []
    for name, steps, P in kat_inputs():
        kat.append({"name": name, "P": P, "steps": hexlist(steps), "L": len(steps),
                    "spline": hexlist(ref.splinify(steps, P))})
    json.dump(kat, open(os.path.join(OUT, "kat_spline.json"), "w"), indent=0)

    # ---- pair KATs: distances and a thresholded edge list
    thr = 1e-6
    rows = synth.rows(11, 96, 8, 10, 5e-3, synth.default_pert(thr, 10))
    rng = np.random.default_rng(5)
    for k in range(12):  # planted pairs a few ulp either side of the threshold
        u = rng.standard_normal(60)
        u /= np.linalg.norm(u)
        rows[48 + k] = rows[k] + u * thr * (1 + (k - 6) * 2e-16)
    ei, ej, ed, _ = ref.all_pairs(rows, thr)
    sample = [(0, 1), (0, 48), (5, 53), (10, 90), (95, 94), (7, 7)]
    pairs = {"thr": thr, "rows": hexlist(rows), "n": 96, "K": 60,
             "l2": [{"a": a, "b": b, "d": ref.compare_l2(rows[a], rows[b]).hex()} for a, b in sample],
             "edges": {"i": ei.tolist(), "j": ej.tolist(), "d": hexlist(ed)}}
    json.dump(pairs, open(os.path.join(OUT, "kat_pairs.json"), "w"), indent=0)

    # ---- config-1-like pipeline through the reference CLIs and the python script
    n, P, L = 576, 10, 12
    off = synth.offsets(1, n, 8, L, L)
    steps = synth.histories(1, n, 8, 1e-4, synth.default_pert(thr, P), off)
    ids = np.arange(n, dtype=np.uint32)
    pdir = os.path.join(OUT, "pipeline_c1")
    shutil.rmtree(pdir, ignore_errors=True)
    os.makedirs(pdir)
    with tempfile.TemporaryDirectory() as td:
        sdir = os.path.join(td, "strains") + "/"
        os.makedirs(sdir)
        os.makedirs(os.path.join(td, "__results"))
        for i in range(n):
            with open(os.path.join(sdir, f"strain_{ids[i]}"), "w") as f:
                for s in steps[int(off[i]):int(off[i + 1])]:
                    f.write(" ".join(repr(float(v)) for v in s) + "\n")
        with open(os.path.join(sdir, "README_not_a_strain_file"), "w") as f:
            f.write("ignored by the CLIs\n")
        subprocess.check_call([ref_binary("mpi_comparison_test"), sdir, str(P), repr(thr)], cwd=td,
                              stdout=subprocess.DEVNULL)
        results = {}
        for i in range(n):
            results[str(i)] = open(os.path.join(td, "__results", f"ID_{i}.txt")).read()
        # the production file names, for the python script
        hdir = os.path.join(td, "macro")
        os.makedirs(hdir)
        for i in range(n):
            shutil.copy(os.path.join(td, "__results", f"ID_{i}.txt"), os.path.join(hdir, f"last.{i}.similar_hist"))
        listing = [nm for nm in os.listdir(hdir)]
        py = subprocess.run([sys.executable, REF_PY, hdir, os.path.join(td, "mapping.csv"), str(n)],
                            capture_output=True, text=True, check=True)
        mapping = open(os.path.join(td, "mapping.csv")).read()
        cmp_out = subprocess.run([ref_binary("compare_all_histories"), sdir, str(P)], capture_output=True,
                                 text=True, check=True).stdout
    cmp_lines = [ln for ln in cmp_out.splitlines() if " vs " in ln]
    json.dump({"n": n, "P": P, "L": L, "thr": thr, "seed": 1, "cluster": 8, "amp": 1e-4,
               "results": results, "glob_order": listing, "mapping_csv": mapping, "python_stdout": py.stdout,
               "compare_all_histories": {"n_vs_lines": len(cmp_lines),
                                         "sha256_sorted_vs_lines": hashlib.sha256("\n".join(sorted(cmp_lines)).encode()).hexdigest(),
                                         "first_sorted": sorted(cmp_lines)[:5]}},
              open(os.path.join(pdir, "reference_outputs.json"), "w"), indent=0)

    # ---- graph reduction cases through the real script
    cases = []
    rng = np.random.default_rng(99)
    for ci, (nn, m) in enumerate(((12, 10), (40, 60), (200, 500), (300, 200), (64, 400))):
        eu = rng.integers(0, nn, size=m)
        ev = rng.integers(0, nn, size=m)
        keep = eu != ev
        eu, ev = eu[keep], ev[keep]
        with tempfile.TemporaryDirectory() as td:
            # one file per source node, like the production layout; both directions present
            byfile = {}
            for a, b in zip(eu.tolist(), ev.tolist()):
                byfile.setdefault(a, []).append((a, b))
                byfile.setdefault(b, []).append((b, a))
            for a, lst in byfile.items():
                with open(os.path.join(td, f"last.{a}.similar_hist"), "w") as f:
                    for (x, y) in lst:
                        f.write(f"{x} {y} 1e-07\n")
            order = [nm for nm in os.listdir(td) if nm.startswith("last.")]
            calls = []
            import glob
            for fname in glob.glob(td + "/last.*.similar_hist"):
                for line in open(fname):
                    c1, c2, _ = line.split()
                    calls.append((int(c1), int(c2)))
            subprocess.run([sys.executable, REF_PY, td, os.path.join(td, "mapping.csv"), str(nn)], check=True,
                           capture_output=True)
            mp = [int(l.split()[1]) for l in open(os.path.join(td, "mapping.csv"))]
        cases.append({"num_gps": nn, "calls": calls, "mapping": mp})
    json.dump(cases, open(os.path.join(OUT, "graph_cases.json"), "w"), indent=0)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
