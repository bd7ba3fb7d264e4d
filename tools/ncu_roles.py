#!/usr/bin/env python
"""Per-role warp-state summary of a k_filter_tc ncu capture (source page): the kernel is warp-specialised, so the
kernel-wide stall chart says little; this groups the samples by the code range of each role (found from the role's
characteristic instruction: UBLKCP = producer, UTCHMMA = MMA issuer, LDTM = epilogue).
  python tools/ncu_roles.py report.ncu-rep [--list ROLE]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(rows[0][1] if len(rows[0]) > 1 else "")
    # role of an instruction = role of the nearest characteristic instruction with the same execution-count class
    marks = []
    for n, r in enumerate(data):
        s = r[idx["Source"]]
        if "UBLKCP" in s: marks.append((n, "producer"))
        elif "UTCHMMA" in s: marks.append((n, "mma"))
        elif "LDTM" in s: marks.append((n, "epilogue"))
    total = sum(int(r[idx["# Samples"]]) for r in data)
    agg = {}
    role_of = []
    for n, r in enumerate(data):
        role = min(marks, key=lambda m: abs(m[0] - n))[1] if marks else "?"
        role_of.append(role)
        a = agg.setdefault(role, {"samples": 0, "instr": 0, "st": {}})
        a["samples"] += int(r[idx["# Samples"]])
        a["instr"] += 1
        for h in stall:
            a["st"][h] = a["st"].get(h, 0) + int(r[idx[h]])
    print("total samples", total)
    for role, a in agg.items():
        top = sorted(a["st"].items(), key=lambda x: -x[1])[:6]
        print("%-9s %5d instr %9d samples (%.1f %%)  %s" % (role, a["instr"], a["samples"], 100.0 * a["samples"] / max(total, 1), top))
    if "--list" in sys.argv:
        want = sys.argv[sys.argv.index("--list") + 1]
        for n, r in enumerate(data):
            if role_of[n] != want or int(r[idx["# Samples"]]) < 200:
                continue
            st = sorted(((h, int(r[idx[h]])) for h in stall if int(r[idx[h]]) > 0), key=lambda x: -x[1])[:2]
            print("%s %7s %10s  %-80s %s" % (r[idx["Address"]][-5:], r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:80], st))


if __name__ == "__main__":
    main()
