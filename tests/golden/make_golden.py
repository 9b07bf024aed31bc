#!/usr/bin/env python3
"""Regenerate tests/golden/* from the reference checkout's own test fixtures.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py [--reference /root/reference]

Only DATA is extracted (sequences, qualities, expected numbers / CIGARs) -- no
reference code is copied.  Sources:
  * tests/resources/pairhmm-testdata.txt            -> pairhmm-testdata.txt (verbatim, 104 rows)
  * tests/smith_waterman_aligner_unit_tests.rs:229-400 -> sw_known_answers.json
  * tests/smith_waterman_aligner_unit_tests.rs:403-1000 (ref_N/read_N/cigar_N) -> sw_long_pairs.json
  * tests/cigar_utils_unit_tests.rs:36-110          -> calculate_cigar_cases.json
"""
import argparse
import json
import os
import re
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))


def sw_known_answers():
    # (reference, read, params[match,mismatch,open,extend], strategy, expected_offset, expected_cigar, source line)
    ORIGINAL_DEFAULT = [3, -1, -4, -3]
    STANDARD_NGS = [25, -50, -110, -6]
    cases = [
        ("AAAGGACTGACTG", "ACTGACTGACTG", ORIGINAL_DEFAULT, "SoftClip", 1, "12M", 230),
        ("AAAGACTACTG", "AACGGACACTG", [50, -100, -220, -12], "SoftClip", 1, "2M2I3M1D4M", 254),
        ("AAAGACTACTG", "AACGGACACTG", [200, -50, -300, -22], "SoftClip", 0, "11M", 261),
        ("AAACCCCC", "CCCCCGGG", ORIGINAL_DEFAULT, "SoftClip", 3, "5M3S", 289),
        ("TGTGTGTGTGTGTGACAGAGAGAGAGAGAGAGAGAGAGAGAGAGA",
         "ACAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGA", STANDARD_NGS, "SoftClip", 14, "31M20S", 306),
        ("AAACCCCC", "CCCCC", ORIGINAL_DEFAULT, "SoftClip", 3, "5M", 380),
        ("AAACCCCC", "CCCCC", ORIGINAL_DEFAULT, "InDel", 0, "3D5M", 381),
        ("AAACCCCC", "CCCCC", ORIGINAL_DEFAULT, "LeadingInDel", 0, "3D5M", 382),
        ("AAACCCCC", "CCCCC", ORIGINAL_DEFAULT, "Ignore", 3, "5M", 383),
    ]
    return [dict(ref=c[0], alt=c[1], params=c[2], strategy=c[3], offset=c[4], cigar=c[5],
                 source="tests/smith_waterman_aligner_unit_tests.rs:%d" % c[6]) for c in cases]


def sw_long_pairs(path):
    src = open(path).read()
    out = []
    for n in (1, 2, 3):
        def vec(name):
            m = re.search(r"let %s_%d: Vec<u8> = vec!\[(.*?)\];" % (name, n), src, re.S)
            return bytes(int(x) for x in re.findall(r"\d+", m.group(1))).decode()
        m = re.search(r"let cigar_%d: Vec<Cigar> = vec!\[(.*?)\];" % n, src, re.S)
        ops = re.findall(r"Cigar::(\w+)\((\d+)\)", m.group(1))
        code = {"Match": "M", "Ins": "I", "Del": "D", "SoftClip": "S"}
        out.append(dict(ref=vec("ref"), alt=vec("read"),
                        listed_cigar="".join("%s%s" % (l, code[o]) for o, l in ops),
                        note="listed_cigar is present in the reference test but its assertion is commented out; "
                             "the test only asserts scalar == gkl"))
    return out


def calculate_cigar_cases(path):
    src = open(path).read()
    cases = re.findall(r'test_compute_cigar\(\s*"([^"]*)",\s*"([^"]*)",\s*"([^"]*)",?\s*\)', src)
    return [dict(ref=a, alt=b, cigar=c) for a, b, c in cases]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    t = os.path.join(args.reference, "tests")
    shutil.copyfile(os.path.join(t, "resources", "pairhmm-testdata.txt"), os.path.join(HERE, "pairhmm-testdata.txt"))
    json.dump(sw_known_answers(), open(os.path.join(HERE, "sw_known_answers.json"), "w"), indent=1)
    json.dump(sw_long_pairs(os.path.join(t, "smith_waterman_aligner_unit_tests.rs")),
              open(os.path.join(HERE, "sw_long_pairs.json"), "w"), indent=1)
    cc = calculate_cigar_cases(os.path.join(t, "cigar_utils_unit_tests.rs"))
    json.dump(cc, open(os.path.join(HERE, "calculate_cigar_cases.json"), "w"), indent=1)
    print("pairhmm rows:", sum(1 for l in open(os.path.join(HERE, "pairhmm-testdata.txt")) if not l.startswith("#")))
    print("calculate_cigar cases:", len(cc))


if __name__ == "__main__":
    main()
