"""CPU tests of the product's host-side CIGAR post-processing (CigarUtils::calculate_cigar minus the alignment itself):
the reference's 114 known answers of tests/cigar_utils_unit_tests.rs:36-110, with the Smith-Waterman result supplied by
the oracle (in production it comes from the GPU batch — see tests/test_gpu_cigar.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lorikeet_b200 import _lib
from lorikeet_b200.context import cigar_to_string
from oracle import lk_oracle as O

CASES = json.load(open(os.path.join(GOLDEN, "calculate_cigar_cases.json")))
NEW_SW_PARAMETERS = [200, -150, -260, -11]


def calculate_cigar_host(ref, alt, params=NEW_SW_PARAMETERS, strategy="InDel"):
    L = _lib.load()
    r = np.frombuffer(ref.encode(), np.uint8) if isinstance(ref, str) else np.asarray(ref, np.uint8)
    a = np.frombuffer(alt.encode(), np.uint8) if isinstance(alt, str) else np.asarray(alt, np.uint8)
    r = np.ascontiguousarray(r); a = np.ascontiguousarray(a)
    max_ops = 128
    ops = np.zeros(max_ops, np.uint32)
    n = C.c_int32(0)
    p = lambda x, t: x.ctypes.data_as(t)
    if L.lk_calculate_cigar_shortcut(p(r, _lib.u8p), r.size, p(a, _lib.u8p), a.size, max_ops, p(ops, _lib.u32p), C.byref(n)):
        return cigar_to_string(ops[:n.value])
    pad = np.full(10, ord("N"), np.uint8)
    pr, pa = np.concatenate([pad, r, pad]), np.concatenate([pad, a, pad])
    sw_ops = np.zeros(4 * (pr.size + pa.size), np.uint32)
    off = C.c_int32(0)
    prm = np.asarray(params, np.int32)
    k = O.lib().lko_sw_align(p(pr, O.u8p), pr.size, p(pa, O.u8p), pa.size, p(prm, O.i32p), O.SW_STRATEGY[strategy],
                             p(sw_ops, O.u32p), sw_ops.size, C.byref(off))
    st = L.lk_calculate_cigar_finish(p(r, _lib.u8p), r.size, p(a, _lib.u8p), a.size, p(sw_ops, _lib.u32p), k, off.value, max_ops,
                                     p(ops, _lib.u32p), C.byref(n))
    assert st == 0
    if n.value == -3:
        return None
    assert n.value >= 0, n.value
    return cigar_to_string(ops[:n.value])


def test_fixture_size():
    assert len(CASES) == 114


@pytest.mark.parametrize("i", range(len(CASES)))
def test_calculate_cigar_known_answers(i):
    c = CASES[i]
    assert calculate_cigar_host(c["ref"], c["alt"]) == c["cigar"], c


def test_cigar_is_consistent_on_random_haplotypes():
    # property: the CIGAR consumes exactly |ref| reference and |alt| alternate bases and replays alt from ref + inserted bases
    import re
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for _ in range(150):
        ref = rng.choice(acgt, int(rng.integers(30, 200)))
        alt = ref.copy()
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(5, len(alt) - 5))
            t = rng.integers(0, 3)
            if t == 0:
                alt[k] = rng.choice(acgt)
            elif t == 1:
                alt = np.concatenate([alt[:k], alt[k + int(rng.integers(1, 9)):]])
            else:
                alt = np.concatenate([alt[:k], rng.choice(acgt, int(rng.integers(1, 9))), alt[k:]])
        cig = calculate_cigar_host(ref, alt)
        assert cig is not None
        r = a = 0
        for n, op in re.findall(r"(\d+)([MID])", cig):
            n = int(n)
            if op == "M":
                r += n; a += n
            elif op == "I":
                a += n
            else:
                r += n
        assert (r, a) == (len(ref), len(alt)), (cig, len(ref), len(alt))
