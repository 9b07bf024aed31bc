"""The SIMD CPU stand-in for the reference's un-vendored gkl AVX PairHMM (oracle/lk_oracle_simd.c; baseline infrastructure,
"ours, not gkl") against the reference's 104 golden rows at the reference's own tolerance
(tests/vector_pair_hmm_unit_tests.rs:63,90: 1e-5 abs) and against the scalar f64 restatement, for every lane width the CPU
here can run, plus the f64 redo on fp32 underflow."""
import ctypes as C

import numpy as np
import pytest

import pairhmm_cases as PC
from conftest import load_pairhmm_golden
from oracle import lk_oracle as O


def _lanes_available():
    L = O.lib()
    L.lko_simd_supported.restype = C.c_int
    return [l for l in (4, 8, 16) if L.lko_simd_supported(l)]


@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_golden_rows_within_reference_tolerance(lanes):
    if lanes not in _lanes_available():
        pytest.skip("this CPU cannot run the %d-lane variant" % lanes)
    rows = load_pairhmm_golden()
    batch = PC.pack_cases([dict(hap=r["hap"], read=r["read"], qual=r["qual"], ins=r["ins"], dele=r["dele"], gcp=r["gcp"]) for r in rows])
    got, redone = O.pairhmm_batch_vector(batch, force_lanes=lanes, want_redone=True)
    exp = np.array([r["expected"] for r in rows])
    assert len(got) == 104 and redone == 0
    assert np.max(np.abs(got - exp)) < 1e-5
    assert np.max(np.abs(got - O.pairhmm_batch(batch)) / np.abs(got)) < 1e-4


def test_regions_layout_n_bases_and_underflow_redo():
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(6, first_region=60)
    b.read_bases[::97] = ord("N")
    b.hap_bases[::131] = ord("N")
    exp = O.pairhmm_batch(b)
    for lanes in _lanes_available():
        got = O.pairhmm_batch_vector(b, force_lanes=lanes)
        assert np.max(np.abs(got - exp) / np.abs(exp)) < 1e-4, lanes
    # an unrelated 600-base read against a 700-base haplotype underflows fp32 (about 1e-250): redone in f64, same answer
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    case = dict(hap=bytes(rng.choice(acgt, 700)), read=bytes(rng.choice(acgt, 600)), qual=np.full(600, 30, np.uint8),
                ins=np.full(600, 45, np.uint8), dele=np.full(600, 45, np.uint8), gcp=np.full(600, 10, np.uint8))
    batch = PC.pack_cases([case])
    got, redone = O.pairhmm_batch_vector(batch, want_redone=True)
    assert redone == 1 and got[0] == O.pairhmm_batch(batch)[0]
    assert O.vector_isa().endswith("fp32")
