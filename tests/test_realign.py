"""CPU tests of the host side of N2 (read -> best-haplotype realignment): the reference's own known answers
(tests/alignment_utils_unit_tests.rs, extracted to tests/golden/realign_cases.json by make_golden.py) for
apply_cigar_to_cigar, append_clipped_elements_from_cigar_to_cigar, read_start_on_reference_haplotype and
create_read_aligned_to_ref (with the Smith-Waterman result supplied by the oracle; in production it comes from the GPU
batch, see tests/test_gpu_realign.py), plus best-allele tie-breaking against a pure-Python restatement."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lorikeet_b200 import _lib
from lorikeet_b200.context import cigar_from_string, cigar_to_string
from oracle import lk_oracle as O

CASES = json.load(open(os.path.join(GOLDEN, "realign_cases.json")))
TO_BEST_HAPLOTYPE = [10, -15, -30, -5]  # ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS, smith_waterman_aligner.rs:22-25
p = lambda x, t: x.ctypes.data_as(t)


def test_fixture_sizes():
    assert {k: len(v) for k, v in CASES.items()} == {"aligned_to_ref": 124, "apply_cigar_to_cigar": 11, "append_clipped": 9,
                                                     "read_start": 6}


@pytest.mark.parametrize("c", CASES["apply_cigar_to_cigar"], ids=lambda c: "%s+%s" % (c["first"], c["second"]))
def test_apply_cigar_to_cigar(c):
    L = _lib.load()
    a, b = cigar_from_string(c["first"]), cigar_from_string(c["second"])
    out = np.zeros(64, np.uint32)
    n = C.c_int32(0)
    assert L.lk_cigar_apply_to_cigar(p(a, _lib.u32p), a.size, p(b, _lib.u32p), b.size, 64, p(out, _lib.u32p), C.byref(n)) == 0
    assert cigar_to_string(out[:n.value]) == c["expected"]


@pytest.mark.parametrize("c", CASES["append_clipped"], ids=lambda c: c["original"])
def test_append_clipped_elements(c):
    L = _lib.load()
    a, b = cigar_from_string(c["shifted"]), cigar_from_string(c["original"])
    out = np.zeros(64, np.uint32)
    n = C.c_int32(0)
    assert L.lk_cigar_append_clipped_elements(p(a, _lib.u32p), a.size, p(b, _lib.u32p), b.size, 64, p(out, _lib.u32p), C.byref(n)) == 0
    assert cigar_to_string(out[:n.value]) == c["expected"]


@pytest.mark.parametrize("c", CASES["read_start"], ids=lambda c: "%s@%d" % (c["cigar"], c["read_start_on_haplotype"]))
def test_read_start_on_reference_haplotype(c):
    L = _lib.load()
    a = cigar_from_string(c["cigar"])
    assert L.lk_read_start_on_reference_haplotype(p(a, _lib.u32p), a.size, c["read_start_on_haplotype"]) == c["expected"]
    assert L.lk_read_start_on_reference_haplotype(p(a, _lib.u32p), a.size, 10 ** 6) == -1  # the reference panics


def realign_host(c):
    """create_read_aligned_to_ref with the oracle's Smith-Waterman in place of the GPU batch."""
    L = _lib.load()
    u8 = lambda s: np.ascontiguousarray(np.frombuffer(s.encode(), np.uint8))
    read, hap, ref = u8(c["read"]), u8(c["hap"]), u8(c["ref"])
    prm = np.asarray(TO_BEST_HAPLOTYPE, np.int32)
    sw_ops = np.zeros(4 * (read.size + hap.size), np.uint32)
    off = C.c_int32(0)
    k = O.lib().lko_sw_align(p(hap, O.u8p), hap.size, p(read, O.u8p), read.size, p(prm, O.i32p), O.SW_STRATEGY["SoftClip"],
                             p(sw_ops, O.u32p), sw_ops.size, C.byref(off))
    oc, hc = cigar_from_string(c["read_cigar"]), cigar_from_string(c["hap_cigar"])
    out = np.zeros(64, np.uint32)
    n, pos = C.c_int32(0), C.c_int64(0)
    st = L.lk_realign_finish(p(read, _lib.u8p), read.size, p(oc, _lib.u32p), oc.size, read.size, p(sw_ops, _lib.u32p), k, off.value,
                             p(hc, _lib.u32p), hc.size, c["hap_start"], p(ref, _lib.u8p), ref.size, c["ref_start"], 64,
                             p(out, _lib.u32p), C.byref(n), C.byref(pos))
    assert st == 0 and n.value >= 0, (st, n.value)
    return pos.value, cigar_to_string(out[:n.value])


@pytest.mark.parametrize("i", range(len(CASES["aligned_to_ref"])))
def test_create_read_aligned_to_ref_known_answers(i):
    c = CASES["aligned_to_ref"][i]
    assert realign_host(c) == (c["expected_pos"], c["expected_cigar"]), c["source"]


def best_allele_py(v, prio, thr=0.2):
    """search_best_allele(.., can_be_reference = true, Some(priorities)), allele_likelihoods.rs:457-554, one read column."""
    A = len(v)
    best, second, best_lk, second_lk = 0, 0, v[0], -np.inf
    for a in range(1, A):
        if v[a] > best_lk:
            second, best, second_lk, best_lk = best, a, best_lk, v[a]
        elif v[a] > second_lk:
            second, second_lk = a, v[a]
    if best_lk - second_lk < thr:
        bp, sp = prio[best], prio[second]
        for a in range(A):
            if a == best or best_lk - v[a] > thr:
                continue
            if prio[a] > bp:
                second, best, sp, bp = best, a, bp, prio[a]
            elif prio[a] > sp:
                second, sp = a, prio[a]
    best_lk = v[best]
    second_lk = v[second] if second != best else -np.inf
    d = best_lk - second_lk
    conf = 0.0 if abs(d) < np.finfo(np.float64).eps else d
    return best, conf, conf > thr


def test_best_alleles_breaking_ties_matches_restatement():
    L = _lib.load()
    rng = np.random.default_rng(5)
    for trial in range(200):
        A, R = int(rng.integers(1, 9)), int(rng.integers(1, 40))
        vals = -rng.random((A, R)) * (3.0 if trial % 2 else 0.3)   # many near-ties in every other trial
        vals[:, rng.integers(0, R)] = vals[0, 0]                  # an exact tie column
        prio = rng.integers(-3, 3, A).astype(np.int32)
        vals = np.ascontiguousarray(vals)
        best = np.zeros(R, np.int32); conf = np.zeros(R); info = np.zeros(R, np.uint8)
        assert L.lk_best_alleles_breaking_ties(p(vals, _lib.f64p), A, R, R, p(prio, _lib.i32p), 0.2, p(best, _lib.i32p),
                                               p(conf, _lib.f64p), p(info, _lib.u8p)) == 0
        for r in range(R):
            b, c, inf = best_allele_py(vals[:, r], prio)
            assert (best[r], conf[r], bool(info[r])) == (b, c, inf), (trial, r)
