"""CPU tests: the oracle's PairHMM restatement against the reference's own fixtures.

Mirrors tests/vector_pair_hmm_unit_tests.rs, tests/pair_hmm_unit_tests.rs and
tests/pair_hmm_model_unit_tests.rs of the reference.
"""
import math

import numpy as np
import pytest

from oracle import lk_oracle as O
import pairhmm_cases as PC


def test_golden_104_rows(pairhmm_golden):
    # tests/vector_pair_hmm_unit_tests.rs:63 — 1e-5 absolute
    assert len(pairhmm_golden) == 104
    worst = 0.0
    for r in pairhmm_golden:
        got = O.pairhmm_forward(r["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"])
        worst = max(worst, abs(got - r["expected"]))
    assert worst < 1e-5, worst


def test_golden_rescaled_equals_plain(pairhmm_golden):
    for r in pairhmm_golden[::7]:
        a = O.pairhmm_forward(r["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"])
        b = O.pairhmm_forward(r["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"], rescaled=True)
        assert abs(a - b) < 1e-10


def test_quality_utils():
    L = O.lib()
    assert L.lko_qual_to_error_prob(30) == pytest.approx(1e-3, rel=1e-12)
    assert L.lko_qual_to_prob(20) == pytest.approx(0.99, rel=1e-12)
    assert L.lko_qual_to_error_prob(0) == 1.0


def test_approximate_log10_sum_log10():
    L = O.lib()
    # exact for integer-phred differences (math_utils.rs:490-493)
    for a, b in [(-1.0, -2.0), (-3.0, -3.0), (-0.6, -4.5), (-4.5, -4.5)]:
        exact = math.log10(10 ** a + 10 ** b)
        assert L.lko_approximate_log10_sum_log10(a, b) == pytest.approx(exact, abs=1e-9)
    assert L.lko_approximate_log10_sum_log10(-math.inf, -2.0) == -2.0
    # beyond MAX_TOLERANCE the smaller term is dropped
    assert L.lko_approximate_log10_sum_log10(-10.0, -1.0) == -1.0


@pytest.mark.parametrize("ins", [5, 10, 20, 30, 45, 60, 123])
@pytest.mark.parametrize("dele", [5, 10, 20, 30, 45, 60, 123])
@pytest.mark.parametrize("gcp", [5, 10, 20])
def test_transition_probs_closed_form(ins, dele, gcp):
    # tests/pair_hmm_model_unit_tests.rs: transitions vs closed form, 1e-9 (1e-6 on MM: Jacobian table)
    L = O.lib()
    t = (O.C.c_double * 6)()
    L.lko_qual_to_trans_probs(t, ins, dele, gcp)
    e = lambda q: 10 ** (-q / 10.0)
    assert t[0] == pytest.approx(1 - e(ins) - e(dele), abs=1e-6)
    assert t[1] == pytest.approx(1 - e(gcp), abs=1e-9)
    assert t[2] == pytest.approx(e(ins), abs=1e-9)
    assert t[3] == pytest.approx(e(gcp), abs=1e-9)
    assert t[4] == pytest.approx(e(dele), abs=1e-9)
    assert t[5] == pytest.approx(e(gcp), abs=1e-9)


def _const(n, v):
    return np.full(n, v, dtype=np.uint8)


def _run_cases(cases):
    batch = PC.pack_cases(cases)
    tri = cases[0]["tristate"]
    assert all(c["tristate"] == tri for c in cases)
    return O.pairhmm_batch(batch, use_tristate=tri)


def test_basic_likelihood_family():
    # make_basic_likelihood_tests: 22 080 analytic cases, 0.2 abs (pair_hmm_unit_tests.rs:99,169)
    cases = PC.basic_likelihood_cases(extensive=True)
    assert len(cases) == 22080
    got = _run_cases(cases)
    exp = np.array([c["expected"] for c in cases])
    assert np.all(got <= 0.0)
    assert np.max(np.abs(got - exp)) <= 0.2


def test_mismatch_sweeps():
    cases = PC.mismatch_sweep_cases()
    got = _run_cases(cases)
    for g, c in zip(got, cases):
        assert abs(g - c["expected"]) <= c["tol"], c["name"]


def test_all_matching_and_simple():
    cases = PC.all_matching_cases()
    got = _run_cases(cases)
    for g, c in zip(got, cases):
        assert g <= 0.0
        if c["expected"] is not None:
            assert abs(g - c["expected"]) <= c["tol"], c["name"]


def test_big_reads_finite():
    cases = PC.big_read_cases()
    got = _run_cases(cases)
    assert np.all(np.isfinite(got)) and np.all(got <= 0.0)


def test_mismatch_scores_lower_and_tristate():
    rl = 30
    read = b"ACGT" * 7 + b"AC"
    hap_ok = b"TT" + read + b"GG"
    mm = bytearray(hap_ok)
    mm[10] = ord("A") if mm[10] != ord("A") else ord("C")
    args = (_const(rl, 30), _const(rl, 45), _const(rl, 45), _const(rl, 10))
    a = O.pairhmm_forward(hap_ok, read, *args)
    b = O.pairhmm_forward(bytes(mm), read, *args)
    c = O.pairhmm_forward(bytes(mm), read, *args, use_tristate=False)
    assert a > b
    assert c > b  # err/1 > err/3
    assert c - b == pytest.approx(math.log10(3.0), abs=1e-2)


def test_n_bases_match_anything():
    rl = 20
    read = b"ACGTACGTACGTACGTACGT"
    hap = b"GG" + read + b"CC"
    hap_n = bytearray(hap)
    hap_n[7] = ord("N")
    read_n = bytearray(read)
    read_n[3] = ord("N")
    args = (_const(rl, 30), _const(rl, 45), _const(rl, 45), _const(rl, 10))
    base = O.pairhmm_forward(hap, read, *args)
    assert O.pairhmm_forward(bytes(hap_n), read, *args) == pytest.approx(base, abs=1e-6)
    assert O.pairhmm_forward(hap, bytes(read_n), *args) == pytest.approx(base, abs=1e-6)


def test_find_first_position_where_haplotypes_differ():
    # tests/pair_hmm_unit_tests.rs:686
    L = O.lib()
    f = lambda a, b: L.lko_find_first_position_where_haplotypes_differ(O._p(O._u8(a), O.u8p), len(a), O._p(O._u8(b), O.u8p), len(b))
    assert f(b"ACGT", b"ACGT") == 4
    assert f(b"ACGT", b"ACCT") == 2
    assert f(b"ACGTAA", b"ACGT") == 4
    assert f(b"TCGT", b"ACGT") == 0


def test_long_read_rescaled_is_finite():
    rng = np.random.default_rng(7)
    H, R = 3000, 2500
    hap = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), H)
    read = hap[100:100 + R].copy()
    flip = rng.random(R) < 0.1
    read[flip] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(flip.sum()))
    q = rng.integers(7, 21, R).astype(np.uint8)
    plain = O.pairhmm_forward(hap, read, q, _const(R, 45), _const(R, 45), _const(R, 10))
    resc = O.pairhmm_forward(hap, read, q, _const(R, 45), _const(R, 45), _const(R, 10), rescaled=True)
    assert math.isfinite(resc) and resc < 0
    if math.isfinite(plain):
        assert resc == pytest.approx(plain, rel=1e-9)
