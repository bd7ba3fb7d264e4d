"""Generate tests/golden/ingest_cases.json from the UNMODIFIED reference reader.

Run in the build container (needs /root/reference -> oracle/_ref):
    python tests/golden/make_golden_ingest.py
Every case is the text of one strain_<ID> file and the steps Strain6D::from_file
(reference headers/strain2spline.h:112-134, i.e. libstdc++ `ifstream >> double`) stored for it,
as hex floats. The cases cover the number grammar corner by corner: signs, leading zeros, bare
points, exponents with and without digits, two points in one token, hex and inf/nan spellings,
overflow and underflow, 17+ digit mantissas, CRLF and blank lines, a short last line.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_cases.json")


def texts():
    rng = np.random.default_rng(77)
    cases = {}
    cases["plain"] = "0 1 2 3 4 5\n0.001 -0.002 0.003 -0.004 0.005 -0.006\n1e-3 2E-3 3e+0 -4E+1 5e0 6e-0\n"
    cases["empty"] = ""
    cases["only_space"] = " \n\t\n"
    cases["short_last_line"] = "1 2 3 4 5 6\n7 8 9 10\n"
    cases["seven_per_line"] = "1 2 3 4 5 6 7\n8 9 10 11 12 13 14\n"
    cases["crlf_tabs_blank"] = "1\t2  3 4 5 6\r\n\r\n\n7 8 9 10 11 12\r\n"
    cases["signs_points"] = "+1 -1 +.5 -.5 5. -5.\n+0 -0 0. .0 -0.0 +00.00\n"
    cases["leading_zeros"] = "007 000.125 -00012.50 0e5 00e-3 0.0e0\n"
    cases["two_points"] = "1.5.25 2 3 4 5\n6 7 8 9 10 11 12\n"
    cases["exp_no_digits"] = "1 2 3 4 5 6\n1e 2 3 4 5 6\n7 8 9 10 11 12\n"
    cases["exp_sign_no_digits"] = "1 2 3 4 5 6\n1e+ 2 3 4 5 6\n"
    cases["double_exp"] = "1e2e3 2 3 4 5 6\n"
    cases["bare_point"] = "1 2 3 4 5 6\n. 2 3 4 5 6\n"
    cases["bare_sign"] = "1 2 3 4 5 6\n- 2 3 4 5 6\n"
    cases["hex"] = "1 2 3 4 5 6\n0x10 2 3 4 5 6\n"
    cases["inf_nan"] = "1 2 3 4 5 6\ninf 2 3 4 5 6\n"
    cases["nan_word"] = "1 2 3 4 5 6\n1 nan 3 4 5 6\n"
    cases["comma"] = "1 2 3 4 5 6\n1,5 2 3 4 5 6\n"
    cases["garbage_suffix"] = "1 2 3 4 5 6abc\n7 8 9 10 11 12\n"
    cases["overflow"] = "1 2 3 4 5 6\n1e400 2 3 4 5 6\n7 8 9 10 11 12\n"
    cases["neg_overflow"] = "1 2 3 4 5 6\n1 -1e999 3 4 5 6\n"
    cases["underflow"] = "1e-400 -1e-400 4.9e-324 2.4e-324 2.5e-324 1e-310\n"
    cases["near_max"] = "1.7976931348623157e308 1.7976931348623158e308 -1.7976931348623157e+308 8.98846567431158e307 1e308 1e-308\n"
    cases["long_mantissa"] = ("0.1000000000000000055511151231257827 123456789012345678901234567890 "
                              "9007199254740993 9007199254740992.5 0.30000000000000004 2.2250738585072011e-308\n")
    cases["many_digits_small"] = "0.000000000000000000000000000001 1000000000000000000000000000000 1e22 1e23 123456789012345e22 123456789012345e23\n"
    cases["halfway"] = "9007199254740993 9007199254740995 1.00000000000000011102230246251565404236316680908203125 1.00000000000000011102230246251565404236316680908203124 1.00000000000000011102230246251565404236316680908203126 5e-324\n"
    cases["exp_huge_digits"] = "1e00000000000000000000005 1e-00000000000000000000005 0e99999999999 1e99999999999 2 3\n"
    lines = []
    for _ in range(40):
        v = rng.standard_normal(6) * 10.0 ** rng.integers(-12, 3, size=6)
        fmt = rng.integers(0, 4)
        if fmt == 0:
            lines.append(" ".join(repr(float(x)) for x in v))
        elif fmt == 1:
            lines.append(" ".join("%.6g" % x for x in v))
        elif fmt == 2:
            lines.append(" ".join("%.16e" % x for x in v))
        else:
            lines.append(" ".join("%.20f" % x for x in v))
    cases["random_formats"] = "\n".join(lines) + "\n"
    return cases


def main():
    ref = Reference()
    out = []
    with tempfile.TemporaryDirectory() as d:
        for k, (name, text) in enumerate(sorted(texts().items())):
            path = os.path.join(d, f"strain_{k}")
            with open(path, "w", newline="") as f:
                f.write(text)
            steps = ref.from_file(path)
            out.append({"name": name, "text": text, "n_steps": int(steps.shape[0]),
                        "steps": [float(x).hex() for x in steps.ravel()]})
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    print(f"wrote {OUT}: {len(out)} cases, {sum(c['n_steps'] for c in out)} steps")


if __name__ == "__main__":
    main()
