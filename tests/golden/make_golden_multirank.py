#!/usr/bin/env python
"""Golden outputs of the reference's multi-rank ring (headers/strain2spline.h:546-614) run with 1, 2, 3 and 4 ranks as
threads: oracle/_ref/multirank_driver_ref = tests/helpers/multirank_driver.cc compiled against the UNMODIFIED reference
header and the thread-backed mpi.h of tests/helpers/mpi_threads (recipe: oracle/Makefile). Writes
tests/golden/multirank_cases.json: [{"args": [...], "stdout": "..."}]."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "multirank_driver_ref")
CASES = [(1, 30, 10, "1e-6", 5), (2, 30, 10, "1e-6", 5), (3, 30, 10, "1e-6", 5), (4, 61, 10, "1e-6", 11), (3, 50, 4, "8e-7", 2),
         (2, 3, 10, "1e-6", 1), (3, 2, 10, "1e-6", 1)]


def main():
    out = []
    for c in CASES:
        args = [str(x) for x in c]
        r = subprocess.run([EXE] + args, capture_output=True, text=True, check=True)
        out.append({"args": args, "stdout": r.stdout})
    with open(os.path.join(HERE, "multirank_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "cases,", sum(o["stdout"].count("\n") for o in out), "lines")


if __name__ == "__main__":
    main()
