"""Pins of the oracle (and of the product's host code) on vectors the REFERENCE's own tests hold, beyond the PairHMM / SW
fixtures: find_number_of_repetitions (tests/variant_context_utils_unit_tests.rs:22-140), the dangling-tail alignments
(tests/read_threading_graph_unit_tests.rs:667-685), the two AlleleLikelihoods post-op properties
(tests/allele_likelihoods_unit_tests.rs:399-446 filter every other read at -100, :725-770 cap at -0.001 symmetric) and the
transition-model grid of tests/pair_hmm_model_unit_tests.rs:16-21.  CPU only; the GPU halves are in tests/test_gpu_pins.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import emul_runner as E
from conftest import GOLDEN
from lorikeet_b200 import _lib
from oracle import lk_oracle as O

REPS = json.load(open(os.path.join(GOLDEN, "find_number_of_repetitions.json")))
TAILS = json.load(open(os.path.join(GOLDEN, "dangling_tail_cases.json")))
u8 = lambda s: np.frombuffer(s.encode(), np.uint8) if s else np.zeros(1, np.uint8)
p8 = lambda a: a.ctypes.data_as(O.u8p)


def test_find_number_of_repetitions_known_answers():
    assert len(REPS) == 20
    Lo, Lp = O.lib(), _lib.load()
    Lo.lko_find_number_of_repetitions.restype = C.c_int
    for c in REPS:
        unit, test = u8(c["unit"]), u8(c["test"])
        got_o = Lo.lko_find_number_of_repetitions(p8(unit), len(c["unit"]), p8(test), len(c["test"]), 1 if c["leading"] else 0)
        got_p = Lp.lk_find_number_of_repetitions(unit.ctypes.data_as(_lib.u8p), len(c["unit"]), test.ctypes.data_as(_lib.u8p),
                                                 len(c["test"]), 1 if c["leading"] else 0)
        assert got_o == c["expected"] and got_p == c["expected"], c
    assert Lp.lk_find_number_of_repetitions(None, 0, None, 0, 1) == -1


def test_dangling_tail_alignments_known_answers_on_the_oracle():
    reached = [c for c in TAILS if c["reached"]]
    assert len(reached) == 8
    for c in reached:
        off, cigar = O.sw_align(c["ref"], c["alt"], c["params"], "LeadingInDel")
        ops = O_cigar_strip_trailing_deletion(cigar)
        assert ops == c["cigar"], c


def O_cigar_strip_trailing_deletion(cigar):  # AlignmentUtils::remove_trailing_deletions, alignment_utils.rs:702-707
    import re
    els = re.findall(r"\d+[MIDS]", cigar)
    return "".join(els[:-1] if els[-1].endswith("D") else els)


def _filter_case(n_alleles, n_reads):
    """make_good_and_bad_likelihoods(.., |r| (r & 1) == 0): even reads keep the initial 0.0, odd reads are -10000 on every
    allele (tests/allele_likelihoods_unit_tests.rs:560-581)"""
    v = np.zeros((n_alleles, n_reads))
    v[:, 1::2] = -10000.0
    return v


@pytest.mark.parametrize("n_alleles,n_reads", [(1, 1), (2, 7), (3, 10), (5, 33), (2, 64)])
def test_filter_poorly_modeled_evidence_keeps_every_other_read(n_alleles, n_reads):
    L = O.lib()
    orig = _filter_case(n_alleles, n_reads)
    v = orig.copy()
    thr = np.full(n_reads, -100.0)
    keep = np.ones(n_reads, np.uint8)
    n_kept = L.lko_filter_poorly_modeled_evidence(v.ctypes.data_as(O.f64p), n_alleles, n_reads, thr.ctypes.data_as(O.f64p), p8(keep))
    assert n_kept == (n_reads + 1) // 2 and list(keep) == [1 - (r & 1) for r in range(n_reads)]
    for r in range(n_kept):
        assert np.array_equal(v[:, r], orig[:, 2 * r])          # new column r is old column 2r (:436-443)
    assert np.all(np.isnan(v[:, n_kept:]))


@pytest.mark.parametrize("seed", range(4))
def test_normalize_caps_the_worst_likelihood(seed):
    L = O.lib()
    rng = np.random.default_rng(seed)
    A, R = int(rng.integers(2, 6)), int(rng.integers(1, 40))
    orig = -np.abs(rng.standard_normal((A, R))) * 3          # fill_two_with_random_likelihoods
    v = orig.copy()
    L.lko_normalize_likelihoods(v.ctypes.data_as(O.f64p), A, R, C.c_double(-0.001), 1, 0)
    exp = np.maximum(orig, orig.max(axis=0, keepdims=True) - 0.001)  # :741-758, best can be alt OR reference (symmetric)
    assert np.array_equal(v, exp)


INS = DEL = [30, 45, 20, 10, 5, 60, 123]
GAP = [10, 20, 5]


def test_transition_model_grid_oracle_and_kernel_tables():
    """pair_hmm_model_unit_tests.rs: quals {30,45,20,10,5,60,123}^2 x gcp {10,20,5} at 1e-9: the oracle's qual_to_trans_probs
    equals the reference's own expectation (quals_to_probs, :123-139, built from match_to_match_prob / qual_to_error_prob), and
    the tables the kernels read (host_model.cpp, through the emulator library's hook) equal the oracle."""
    Lo, Le = O.lib(), E.lib()
    Lo.lko_match_to_match_prob.restype = C.c_double
    Lo.lko_qual_to_error_prob.restype = C.c_double
    Lo.lko_qual_to_prob.restype = C.c_double
    Le.emul_model_trans_probs.restype = None
    for ins in INS:
        for dele in DEL:
            for gcp in GAP:
                o = (C.c_double * 6)()
                Lo.lko_qual_to_trans_probs(o, ins, dele, gcp)
                k = (C.c_double * 6)()
                Le.emul_model_trans_probs(ins, dele, gcp, k)
                exp = [Lo.lko_match_to_match_prob(ins, dele), Lo.lko_qual_to_prob(gcp), Lo.lko_qual_to_error_prob(ins),
                       Lo.lko_qual_to_error_prob(gcp), Lo.lko_qual_to_error_prob(dele), Lo.lko_qual_to_error_prob(gcp)]
                for i in range(6):
                    assert abs(o[i] - exp[i]) <= 1e-9 * max(1.0, abs(exp[i])), (ins, dele, gcp, i)
                    assert abs(k[i] - o[i]) <= 1e-9 * max(1.0, abs(o[i])), (ins, dele, gcp, i, k[i], o[i])
                # and the closed forms the table approximates (the Jacobian-log table is good to ~1e-5 relative)
                assert abs(o[2] - 10 ** (-ins / 10)) < 1e-12 and abs(o[4] - 10 ** (-dele / 10)) < 1e-12
                assert abs(o[0] - max(0.0, 1 - 10 ** (-ins / 10) - 10 ** (-dele / 10))) < 2e-4
