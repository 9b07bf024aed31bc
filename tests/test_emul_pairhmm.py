"""CPU tests of the CUDA kernels' warp routines through the lock-step emulator (tests/emul).

These run the same templated code the GPU runs (pairhmm_core.cuh) so lane/strip/skew index bugs
are caught without a GPU.  The authoritative parity tests are the -m gpu ones."""
import numpy as np
import pytest

import emul_runner as E
import pairhmm_cases as PC
from conftest import load_pairhmm_golden
from oracle import lk_oracle as O


def _golden_batch(rows):
    cases = [dict(hap=r["hap"], read=r["read"], qual=r["qual"], ins=r["ins"], dele=r["dele"], gcp=r["gcp"]) for r in rows]
    return PC.pack_cases(cases)


@pytest.mark.parametrize("force_k", [0, 4, 10, 15, 19])
def test_golden_rows_fast_path(force_k):
    rows = load_pairhmm_golden()[::4]
    got, flags, n_resc = E.pairhmm_batch(_golden_batch(rows), force_k=force_k)
    exp = np.array([r["expected"] for r in rows])
    assert n_resc == 0 and np.all(flags == 0)
    assert np.max(np.abs(got - exp)) < 1e-5  # tests/vector_pair_hmm_unit_tests.rs:63


def test_golden_rows_f64_pass():
    rows = load_pairhmm_golden()[::8]
    batch = _golden_batch(rows)
    got, flags, n_resc = E.pairhmm_batch(batch, force_f64=True)
    exp = O.pairhmm_batch(batch)
    assert n_resc == len(rows) and np.all(flags == 1)
    assert np.max(np.abs(got - exp)) < 1e-11


def test_region_layout_and_packing():
    # 3 regions with ragged read lengths / haplotype counts: out[r*A + h], several reads per warp
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    rrs, rhs, roff, hoff = [0], [0], [0], [0]
    rb, rq, ri, rd, hb = [], [], [], [], []
    for g in range(3):
        base = rng.choice(acgt, 120 + 20 * g)
        for _ in range(2 + g):
            h = base.copy()
            h[rng.integers(0, len(h), 2)] = rng.choice(acgt, 2)
            hb.append(h); hoff.append(hoff[-1] + len(h))
        rhs.append(rhs[-1] + 2 + g)
        for _ in range(7):
            n = int(rng.integers(5, 100))
            s = int(rng.integers(0, len(base) - n))
            r = base[s:s + n].copy()
            if n > 10:
                r[rng.integers(0, n)] = ord("N")
            rb.append(r); rq.append(rng.integers(6, 41, n).astype(np.uint8))
            ri.append(rng.integers(30, 46, n).astype(np.uint8)); rd.append(rng.integers(30, 46, n).astype(np.uint8))
            roff.append(roff[-1] + n)
        rrs.append(rrs[-1] + 7)
    batch = dict(region_read_start=np.array(rrs), region_hap_start=np.array(rhs), read_off=np.array(roff),
                 hap_off=np.array(hoff), read_bases=np.concatenate(rb), read_qual=np.concatenate(rq),
                 read_ins=np.concatenate(ri), read_del=np.concatenate(rd),
                 read_gcp=np.full(roff[-1], 10, np.uint8), hap_bases=np.concatenate(hb))
    exp = O.pairhmm_batch(batch)
    for fk in (0, 6, 12):
        got, flags, n_resc = E.pairhmm_batch(batch, force_k=fk)
        assert n_resc == 0
        assert np.max(np.abs(got - exp) / np.abs(exp)) < 1e-4


@pytest.mark.parametrize("gcp", [10, 12])
def test_small_launches_split_along_the_haplotype_axis(monkeypatch, gcp):
    # host_prep cuts the warp tasks of a small launch into haplotype slices (>= 4 haplotypes each) so a few regions still
    # fill the GPU; every pair must come out bit-for-bit as from the whole-region tasks, in the same out[r*A + h] slots.
    # gcp = 10 runs the SPLIT instantiation; any other gcp runs the split task list through the whole-region routine
    # (each task then sweeps all haplotypes: redundant, never wrong).
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    rrs, rhs, roff, hoff = [0], [0], [0], [0]
    rb, rq, ri, rd, hb = [], [], [], [], []
    for g, n_haps in enumerate((13, 3, 9, 16)):
        base = rng.choice(acgt, 90 + 10 * g)
        for _ in range(n_haps):
            h = base.copy()
            h[rng.integers(0, len(h), 2)] = rng.choice(acgt, 2)
            hb.append(h); hoff.append(hoff[-1] + len(h))
        rhs.append(rhs[-1] + n_haps)
        for _ in range(5):
            n = int(rng.integers(20, 80))
            s0 = int(rng.integers(0, len(base) - n))
            rb.append(base[s0:s0 + n].copy()); rq.append(rng.integers(6, 41, n).astype(np.uint8))
            ri.append(np.full(n, 45, np.uint8)); rd.append(np.full(n, 45, np.uint8))
            roff.append(roff[-1] + n)
        rrs.append(rrs[-1] + 5)
    batch = dict(region_read_start=np.array(rrs), region_hap_start=np.array(rhs), read_off=np.array(roff),
                 hap_off=np.array(hoff), read_bases=np.concatenate(rb), read_qual=np.concatenate(rq),
                 read_ins=np.concatenate(ri), read_del=np.concatenate(rd),
                 read_gcp=np.full(roff[-1], gcp, np.uint8), hap_bases=np.concatenate(hb))
    # the split really happens: more tasks than the whole-region packing (one launch per K class, each cut on its own)
    assert E.count_fast_tasks(batch) < E.count_fast_tasks(batch, 7) <= E.count_fast_tasks(batch, 1000)
    monkeypatch.delenv("LK_EMUL_SPLIT_TASKS", raising=False)
    whole, _, _ = E.pairhmm_batch(batch)
    exp = O.pairhmm_batch(batch)
    assert np.max(np.abs(whole - exp) / np.abs(exp)) < 1e-4
    for target in ("7", "1000"):
        monkeypatch.setenv("LK_EMUL_SPLIT_TASKS", target)
        got, flags, n_resc = E.pairhmm_batch(batch)
        assert n_resc == 0 and np.array_equal(got, whole)


def test_analytic_families_subset():
    cases = PC.basic_likelihood_cases(extensive=True)[::97] + PC.mismatch_sweep_cases()[::5] + PC.all_matching_cases()
    batch = PC.pack_cases(cases)
    exp = O.pairhmm_batch(batch, use_tristate=False)
    got, flags, n_resc = E.pairhmm_batch(batch, use_tristate=False)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - exp) / np.maximum(1e-3, np.abs(exp))) < 1e-4


def test_variable_gcp_and_exotic_bases():
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = rng.choice(acgt, 90)
    read = hap[10:70].copy()
    n = len(read)
    gcp = rng.integers(5, 20, n).astype(np.uint8)
    c1 = dict(hap=hap, read=read, qual=np.full(n, 30, np.uint8), ins=np.full(n, 40, np.uint8),
              dele=np.full(n, 42, np.uint8), gcp=gcp)
    hap2 = hap.copy(); hap2[30] = ord("X"); read2 = read.copy(); read2[20] = ord("X")
    c2 = dict(c1, hap=hap2, read=read2)
    batch = PC.pack_cases([c1, c2])
    exp = O.pairhmm_batch(batch)
    got, flags, n_resc = E.pairhmm_batch(batch)
    assert n_resc == 1 and flags[0] == 0 and flags[1] == 1  # exotic bytes go to the byte-exact f64 pass
    assert abs(got[0] - exp[0]) / abs(exp[0]) < 1e-4
    assert abs(got[1] - exp[1]) < 1e-10


def test_underflow_is_rescued_and_long_read_bands():
    rng = np.random.default_rng(2)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    # (a) unrelated read: fp32 underflows -> f64 pass
    hap = rng.choice(acgt, 200)
    read = rng.choice(acgt, 150)
    n = len(read)
    c1 = dict(hap=hap, read=read, qual=np.full(n, 40, np.uint8), ins=np.full(n, 45, np.uint8),
              dele=np.full(n, 45, np.uint8), gcp=np.full(n, 10, np.uint8))
    # (b) read longer than the fast path's 32*20 rows -> band kernel (fp32), or the f64 pass when forced
    hap_l = rng.choice(acgt, 900)
    read_l = hap_l[50:50 + 700].copy()
    m = len(read_l)
    c2 = dict(hap=hap_l, read=read_l, qual=rng.integers(10, 30, m).astype(np.uint8), ins=np.full(m, 45, np.uint8),
              dele=np.full(m, 45, np.uint8), gcp=np.full(m, 10, np.uint8))
    batch = PC.pack_cases([c1, c2])
    exp = O.pairhmm_batch(batch)
    exp_resc = O.pairhmm_batch(batch, rescaled=True)
    got, flags, n_resc = E.pairhmm_batch(batch)
    assert n_resc == 1 and list(flags) == [1, 0]   # (b) now runs in the fp32 band kernel
    assert abs(got[0] - exp[0]) < 1e-9
    assert abs(got[1] - exp_resc[1]) / abs(exp_resc[1]) < 1e-4
    got, flags, n_resc = E.pairhmm_batch(batch, force_f64=True)  # the f64 pass still handles several row bands
    assert abs(got[1] - exp_resc[1]) < 1e-8


def _noisy_long_case(rng, R, H, qlo=7, qhi=21, indel=0.03):
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = rng.choice(acgt, H)
    out, pos = [], int(rng.integers(0, H - R - 60))
    while len(out) < R:
        u = rng.random()
        if u < indel / 2:
            out.append(rng.choice(acgt))
        else:
            if u < indel:
                pos += 1
            out.append(hap[pos]); pos += 1
    read = np.array(out, dtype=np.uint8)
    q = rng.integers(qlo, qhi, R).astype(np.uint8)
    flip = rng.random(R) < 10 ** (-q.astype(np.float64) / 10.0)
    read[flip] = rng.choice(acgt, int(flip.sum()))
    return dict(hap=hap, read=read, qual=q, ins=np.full(R, 45, np.uint8), dele=np.full(R, 45, np.uint8), gcp=np.full(R, 10, np.uint8))


def test_band_path_many_bands_noisy_reads():
    # K = 4 -> bands of 128 rows: 700-1100 row reads cross 6-9 bands; ONT-like noise forces the exponent schedule to work
    rng = np.random.default_rng(21)
    cases = [_noisy_long_case(rng, 700, 900), _noisy_long_case(rng, 1100, 1300), _noisy_long_case(rng, 513, 800, 30, 41, 0.0),
             _noisy_long_case(rng, 128, 300), _noisy_long_case(rng, 129, 300)]
    batch = PC.pack_cases(cases)
    exp = O.pairhmm_batch(batch, rescaled=True)
    got, flags, n_resc = E.pairhmm_batch(batch, force_k=-4)
    assert n_resc == 0 and np.all(flags == 0)
    assert np.max(np.abs(got - exp) / np.abs(exp)) < 1e-4


def test_band_path_default_k_and_hopeless_pair():
    rng = np.random.default_rng(22)
    ok = _noisy_long_case(rng, 900, 1200)
    bad = dict(ok, read=rng.choice(np.frombuffer(b"ACGT", np.uint8), 900), qual=np.full(900, 40, np.uint8))  # unrelated read
    batch = PC.pack_cases([ok, bad])
    exp = O.pairhmm_batch(batch, rescaled=True)
    got, flags, n_resc = E.pairhmm_batch(batch)  # R = 900 > 32*20: band kernel with its production K
    assert n_resc == 0 and np.all(flags == 0)    # block floating point also copes with the ~1e-1700 of an unrelated read
    assert np.max(np.abs(got - exp) / np.abs(exp)) < 1e-4
