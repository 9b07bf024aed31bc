"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/*.h declares,
rejects bad arguments without a GPU, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from lorikeet_b200 import _lib


def _declared_symbols(hdr):
    src = open(os.path.join(ROOT, "include", hdr)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lk_[a-z0-9_]+)\s*\(", src)))


def _exported_symbols(so_path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], check=True, capture_output=True, text=True).stdout
    return sorted(t[2] for t in (ln.split() for ln in out.splitlines()) if len(t) == 3 and t[1] == "T" and t[2].startswith("lk_"))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared_symbols("lorikeet_b200.h")
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    assert set(_lib.EXPORTED_SYMBOLS) <= set(declared)
    # and nothing is exported behind the header's back
    assert set(_exported_symbols(_lib.SO_PATH)) <= set(declared)


def test_generator_library_is_separate_and_complete():
    from lorikeet_b200 import synth
    S = C.CDLL(synth.SYNTH_SO_PATH)
    for name in _declared_symbols("lorikeet_b200_synth.h"):
        assert hasattr(S, name), "missing export: " + name
    # the generators are not part of the product library, and the generator library has no CUDA dependency
    assert not [n for n in _exported_symbols(_lib.SO_PATH) if n.startswith("lk_synth")]
    import subprocess
    needed = subprocess.run(["readelf", "-d", synth.SYNTH_SO_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in needed.lower()


def test_version_and_dims_without_gpu():
    L = _lib.load()
    assert b"sm_100a" in L.lk_version()
    from lorikeet_b200 import PairHmmBatch
    b = PairHmmBatch([0, 2], [0, 3], [0, 5, 12], b"ACGTAACGTACG", bytes([30] * 12), bytes([45] * 12), bytes([45] * 12),
                     [0, 10, 20, 31], b"A" * 31)
    s = b.as_struct()
    n_pairs, n_cells = C.c_int64(), C.c_int64()
    assert L.lk_pairhmm_batch_dims(C.byref(s), C.byref(n_pairs), C.byref(n_cells)) == 0
    assert n_pairs.value == 6 == b.n_pairs
    assert n_cells.value == 12 * 31 == b.n_cells
    bad = PairHmmBatch([0, 2], [0, 3], [0, 5, 5], b"ACGTA", bytes([30] * 5), bytes([45] * 5), bytes([45] * 5),
                       [0, 10, 20, 31], b"A" * 31)  # empty read
    s = bad.as_struct()
    assert L.lk_pairhmm_batch_dims(C.byref(s), C.byref(n_pairs), C.byref(n_cells)) == _lib.LK_ERR_INVALID


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import lorikeet_b200 as lk
    with pytest.raises(lk.LorikeetError) as ei:
        lk.Context(0)
    assert ei.value.status == _lib.LK_ERR_CUDA


def test_engine_modify_read_qualities_matches_oracle():
    # host-side pre-op (runs without a GPU): PCR indel model + quality caps == oracle restatement
    from oracle import lk_oracle as O
    L = _lib.load()
    OL = O.lib()
    rng = np.random.default_rng(0)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads, offs = [], [0]
    for i in range(60):
        n = int(rng.integers(1, 160))
        r = rng.choice(acgt, n)
        if i % 3 == 0 and n > 30:   # plant tandem repeats / homopolymers
            k = int(rng.integers(0, n - 25))
            unit = rng.choice(acgt, int(rng.integers(1, 5)))
            rep = np.tile(unit, 25)[:24]
            r[k:k + 24] = rep
        reads.append(r); offs.append(offs[-1] + n)
    bases = np.concatenate(reads)
    n_total = len(bases)
    for pcr in (0, 1, 2, 3):
        for cap_off in (0, 1):
            qual = rng.integers(2, 42, n_total).astype(np.uint8)
            ins = rng.integers(2, 46, n_total).astype(np.uint8)
            dele = rng.integers(2, 46, n_total).astype(np.uint8)
            mapq = rng.integers(0, 61, len(reads)).astype(np.uint8)
            q1, i1, d1 = qual.copy(), ins.copy(), dele.copy()
            args = _lib.EngineArgs(10, -4.5, pcr, 18, 0, 1.0, 0.02, 1, cap_off, 0)
            off = np.array(offs, dtype=np.int64)
            p = lambda a, t: a.ctypes.data_as(t)
            st = L.lk_engine_modify_read_qualities(C.byref(args), len(reads), p(off, _lib.i64p), p(bases, _lib.u8p),
                                                   p(q1, _lib.u8p), p(i1, _lib.u8p), p(d1, _lib.u8p), p(mapq, _lib.u8p), 2)
            assert st == 0
            q2, i2, d2 = qual.copy(), ins.copy(), dele.copy()
            for r in range(len(reads)):
                a, b = offs[r], offs[r + 1]
                bb = np.ascontiguousarray(bases[a:b]); ii = np.ascontiguousarray(i2[a:b]); dd = np.ascontiguousarray(d2[a:b]); qq = np.ascontiguousarray(q2[a:b])
                OL.lko_apply_pcr_error_model(p(bb, O.u8p), b - a, p(ii, O.u8p), p(dd, O.u8p), pcr)
                OL.lko_cap_minimum_read_qualities(p(qq, O.u8p), p(ii, O.u8p), p(dd, O.u8p), b - a, int(mapq[r]), 18, cap_off)
                i2[a:b], d2[a:b], q2[a:b] = ii, dd, qq
            assert np.array_equal(q1, q2) and np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_pcr_cache_values_match_survey_table():
    # cache[n] for Conservative (rate 3): 40 39 39 39 39 39 39 38 38 38 38 37 37 ...; Hostile clamps to 6 at 12
    from oracle import lk_oracle as O
    OL = O.lib()
    cons = [OL.lko_pcr_error_model_adjusted_qual(n, 3.0) for n in range(13)]
    assert cons == [40, 39, 39, 39, 39, 39, 39, 38, 38, 38, 38, 37, 37]
    host = [OL.lko_pcr_error_model_adjusted_qual(n, 1.0) for n in range(13)]
    assert host == [40, 39, 39, 38, 37, 36, 34, 31, 28, 23, 16, 7, 6]
    assert OL.lko_pcr_error_model_adjusted_qual(30, 3.0) == 16 and OL.lko_pcr_error_model_adjusted_qual(100, 1.0) == 6


def test_reference_arm_does_not_map_the_product_library():
    """bench.py --impl reference builds its inputs through the generator library only: the process must never map
    liblorikeet_b200.so (the driver records the libraries each arm loads)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from lorikeet_b200 import synth\n"
            "from oracle import lk_oracle as O\n"
            "b = synth.pairhmm_regions(2)\n"
            "O.pairhmm_batch(b, n_threads=2); O.pairhmm_batch_vector(b, n_threads=2)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'liblorikeet_b200_synth.so' in maps and 'liblk_oracle' in maps\n"
            "assert 'liblorikeet_b200.so' not in maps, 'product library mapped'\n" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_pool_has_no_cpu_fallback_either():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.load()
    h = C.c_void_p()
    assert L.lk_pool_create(None, 0, C.byref(h)) == _lib.LK_ERR_CUDA and not h.value
    from lorikeet_b200.engine import DevicePool
    with pytest.raises(_lib.LorikeetError):
        DevicePool()
