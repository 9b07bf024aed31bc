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
  * tests/alignment_utils_unit_tests.rs:155-340 (create_read_aligned_to_ref known answers), :840-878
    (apply_cigar_to_cigar), :902-916 (append_clipped_elements...), :1017-1024 (read_start_on_reference_haplotype)
                                                     -> realign_cases.json
  * tests/variant_context_utils_unit_tests.rs:22-140 (find_number_of_repetitions) -> find_number_of_repetitions.json
  * tests/read_threading_graph_unit_tests.rs:667-685 (make_dangling_tails_data)  -> dangling_tail_cases.json
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


def realign_cases(path):
    """Known answers of the reference's create_read_aligned_to_ref tests.  The data provider
    (make_read_aligned_to_ref_data, tests/alignment_utils_unit_tests.rs:155-340) builds most cases in loops from the
    literals below; the loops are replayed here, the three long literal examples are pulled out of the file."""
    src = open(path).read()
    out = {"aligned_to_ref": [], "apply_cigar_to_cigar": [], "append_clipped": [], "read_start": []}
    hap = re.search(r'let hap_bases = "([ACGT]+)";', src).group(1)
    allm = "%dM" % len(hap)

    def case(read, hap_bases, hap_cigar, hap_start, ref_bases, ref_start, pos, cigar, line):
        out["aligned_to_ref"].append(dict(read=read, hap=hap_bases, hap_cigar=hap_cigar, hap_start=hap_start, ref=ref_bases,
                                          ref_start=ref_start, expected_pos=pos, expected_cigar=cigar, read_cigar="10M",
                                          source="tests/alignment_utils_unit_tests.rs:%d" % line))
    for i in range(-1, len(hap)):   # :160-180 one substitution to 'A' anywhere: still all-M
        b = list(hap)
        if i != -1:
            b[i] = "A"
        case("".join(b), hap, allm, 0, hap, 10, 10, allm, 160)
    for pad in range(1, 10):        # :183-199 insertions at the front
        case("N" * pad + hap, hap, allm, 0, hap, 10, 10, "%dI%s" % (pad, allm), 183)
    for pad in range(1, 10):        # :202-217 insertions at the back
        case(hap + "N" * pad, hap, allm, 0, hap, 10, 10, "%s%dI" % (allm, pad), 202)
    for ref_start in range(1, 10):  # :220-238 ref_start and hap_start are respected
        for hap_start in range(ref_start, 10 + ref_start):
            case(hap, hap, allm, hap_start, hap, ref_start, ref_start + hap_start, allm, 220)
    # :241-262 SW does not left-align: 31M6D211M must come out as 28M6D214M
    m = re.search(r'let hap = "([ACGT]+)";\s*let hap_cigar = format!\("\{\}M", hap.len\(\)\);\s*let read_bases = "([ACGT]+)";', src)
    case(m.group(2), m.group(1), "%dM" % len(m.group(1)), 500, m.group(1), 10130100, 10130740, "28M6D214M", 241)
    # :265-296 left-align generates a leading deletion; :299-335 haplotype with an indel relative to the reference
    for m in re.finditer(r'let reference = "([ACGT]+)";.*?let haplotype = "([ACGT]+)";.*?let read_bases = "([ACGT]+)";\s*'
                         r'let hap_cigar = ([^;]+);.*?let ref_start = (\d+);\s*let hap_start = (\d+);\s*'
                         r'let alignment_offset = (\d+);\s*let good_cigar = "(\w+)";', src, re.S):
        ref, h, rd, hc, rs, hs, ao, good = m.groups()
        hc = hc.strip()
        hap_cigar = "%dM" % len(h) if hc.startswith("format!") else hc.strip('"')
        case(rd, h, hap_cigar, int(hs), ref, int(rs), int(rs) + int(hs) + int(ao), good, src[:m.start()].count("\n") + 1)
    for i in range(1, 5):  # :840-846
        out["apply_cigar_to_cigar"].append(dict(first="%dM" % i, second="%dM" % i, expected="%dM" % i))
    for a, b, c in re.findall(r'test_apply_cigar_to_cigar\("(\w+)", "(\w+)", "(\w+)"\);', src):
        out["apply_cigar_to_cigar"].append(dict(first=a, second=b, expected=c))
    for a, b, c in re.findall(r'test_append_clipped_elements_from_original_cigar\(\s*"(\w+)",\s*"(\w+)",\s*"(\w+)",?\s*\)', src):
        out["append_clipped"].append(dict(original=a, shifted=b, expected=c))
    for a, b, c in re.findall(r'test_read_start_on_reference_haplotype\("(\w+)", (\d+), (\d+)\);', src):
        out["read_start"].append(dict(cigar=a, read_start_on_haplotype=int(b), expected=int(c)))
    return out


def find_number_of_repetitions_cases(path):
    """assert_eq!(VariantContextUtils::find_number_of_repetitions("AT".as_bytes(), "GATAT".as_bytes(), false), 2)"""
    src = open(path).read()
    pat = re.compile(r'find_number_of_repetitions\(\s*"(\w*)"\.as_bytes\(\),\s*"(\w*)"\.as_bytes\(\),\s*(true|false),?\s*\),\s*(\d+)', re.S)
    return [dict(unit=u, test=t, leading=(l == "true"), expected=int(e)) for u, t, l, e in pat.findall(src)]


def dangling_tail_cases(path):
    """test_dangling_tails(ref_end, alt_end, cigar, cigar_is_good, ...): the graph hands SmithWatermanAligner::align the path
    bases from the last common vertex on, i.e. the last base of the common prefix ('T') + each end ("add 1M to the expected
    CIGAR because it includes the previous (common) base too"), STANDARD_NGS, LeadingInDel, then remove_trailing_deletions
    (read_threading_graph.rs:1408-1421).  Only cases whose alignment the reference test actually reaches are kept: for a
    dangling path shorter than min_tail_path_length + 1 = 5 vertices the graph returns None before aligning, and the test then
    only asserts !cigar_is_good, so the listed CIGAR of those two rows ("C", "CA") is never compared."""
    src = open(path).read()
    prefix = re.search(r'let common_prefix = "([ACGT]+)";', src).group(1)
    seen, out = set(), []
    for r, a, c, good in re.findall(r'test_dangling_tails\("(\w+)", "(\w+)", "(\w+)", (true|false),', src):
        if (r, a) in seen:
            continue
        seen.add((r, a))
        out.append(dict(ref=prefix[-1] + r, alt=prefix[-1] + a, cigar=c, params=[25, -50, -110, -6], reached=len(a) >= 4,
                        cigar_is_good=(good == "true")))
    return out


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
    rc = realign_cases(os.path.join(t, "alignment_utils_unit_tests.rs"))
    json.dump(rc, open(os.path.join(HERE, "realign_cases.json"), "w"), indent=1)
    fr = find_number_of_repetitions_cases(os.path.join(t, "variant_context_utils_unit_tests.rs"))
    json.dump(fr, open(os.path.join(HERE, "find_number_of_repetitions.json"), "w"), indent=1)
    dt = dangling_tail_cases(os.path.join(t, "read_threading_graph_unit_tests.rs"))
    json.dump(dt, open(os.path.join(HERE, "dangling_tail_cases.json"), "w"), indent=1)
    print("find_number_of_repetitions cases:", len(fr), " dangling tail cases:", len(dt))
    print("realign cases:", {k: len(v) for k, v in rc.items()})
    print("pairhmm rows:", sum(1 for l in open(os.path.join(HERE, "pairhmm-testdata.txt")) if not l.startswith("#")))
    print("calculate_cigar cases:", len(cc))


if __name__ == "__main__":
    main()
