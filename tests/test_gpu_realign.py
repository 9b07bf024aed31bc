"""GPU parity of N2: lk_realign_reads_batch (best haplotype on the host, ONE Smith-Waterman batch on the GPU, CIGAR
composition on the host) against the reference's known answers for create_read_aligned_to_ref and against the
host pieces driven by the oracle's Smith-Waterman on random regions."""
import json
import os

import numpy as np
import pytest

import lorikeet_b200 as lk
from conftest import GOLDEN
from test_realign import realign_host

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(GOLDEN, "realign_cases.json")))["aligned_to_ref"]


def test_known_answers_as_one_batch():
    # every golden case becomes a region with two haplotypes: [reference haplotype, the case's haplotype]; the likelihood
    # matrix makes the second the best allele by a wide margin
    rrs, rhs, vals, haps, hcig, hstart, reads, rcig, rstart = [0], [0], [], [], [], [], [], [], []
    for c in CASES:
        haps += [c["ref"], c["hap"]]
        hcig += ["%dM" % len(c["ref"]), c["hap_cigar"]]
        hstart += [0, c["hap_start"]]
        reads.append(c["read"]); rcig.append(c["read_cigar"]); rstart.append(c["ref_start"])
        vals += [-10.0, -1.0]
        rrs.append(len(reads)); rhs.append(len(haps))
    b = lk.RealignBatch(rrs, rhs, vals, haps, hcig, hstart, reads, rcig, rstart)
    with lk.Context(0) as ctx:
        best, info, cigars, pos = ctx.realign_reads_batch(b)
    assert np.all(best == 1) and np.all(info == 1)
    for i, c in enumerate(CASES):
        assert (int(pos[i]), cigars[i]) == (c["expected_pos"], c["expected_cigar"]), c["source"]


def test_random_regions_match_host_pieces_with_oracle_sw():
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    rrs, rhs, vals, haps, hcig, hstart, reads, rcig, rstart, expect = [0], [0], [], [], [], [], [], [], [], []
    for g in range(24):
        ref = rng.choice(acgt, 220)
        A = int(rng.integers(2, 6))
        region_haps, region_cigs = [ref], ["220M"]
        for a in range(1, A):
            k = int(rng.integers(40, 180))
            if a % 2:   # deletion of d bases
                d = int(rng.integers(1, 7))
                region_haps.append(np.concatenate([ref[:k], ref[k + d:]])); region_cigs.append("%dM%dD%dM" % (k, d, 220 - k - d))
            else:       # insertion
                d = int(rng.integers(1, 7))
                region_haps.append(np.concatenate([ref[:k], rng.choice(acgt, d), ref[k:]])); region_cigs.append("%dM%dI%dM" % (k, d, 220 - k))
        R = int(rng.integers(3, 12))
        V = -rng.random((A, R)) * 5
        region_start = int(rng.integers(1, 10 ** 6))
        for r in range(R):
            a = int(np.argmax(V[:, r]))
            V[a, r] += 1.0  # informative margin, so the best allele is unambiguous
            h = region_haps[a]
            s = int(rng.integers(0, len(h) - 80))
            rd = h[s:s + 80].copy()
            rd[int(rng.integers(0, 80))] = rng.choice(acgt)
            reads.append(rd); rcig.append("80M")
            expect.append(realign_host(dict(read=rd.tobytes().decode(), hap=h.tobytes().decode(), ref=ref.tobytes().decode(),
                                            hap_cigar=region_cigs[a], read_cigar="80M", hap_start=0, ref_start=region_start)))
        haps += region_haps; hcig += region_cigs; hstart += [0] * A
        vals.append(V.ravel()); rstart.append(region_start)
        rrs.append(len(reads)); rhs.append(len(haps))
    b = lk.RealignBatch(rrs, rhs, np.concatenate(vals), haps, hcig, hstart, reads, rcig, rstart)
    with lk.Context(0) as ctx:
        best, info, cigars, pos = ctx.realign_reads_batch(b)
    for i, (epos, ecig) in enumerate(expect):
        assert (int(pos[i]), cigars[i]) == (epos, ecig), i


def test_large_batch_multithreaded_host_path_matches_small_batch():
    # above 2048 reads the gather and the CIGAR composition of lk_realign_reads_batch run on several host threads: tiling
    # one small set of regions many times must give, tile by tile, exactly what the small (single-threaded) batch gives
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    rrs, rhs, vals, haps, hcig, hstart, reads, rcig, rstart = [0], [0], [], [], [], [], [], [], []
    for g in range(12):
        ref = rng.choice(acgt, 200)
        k, d = int(rng.integers(40, 150)), int(rng.integers(1, 6))
        region_haps = [ref, np.concatenate([ref[:k], ref[k + d:]]), np.concatenate([ref[:k], rng.choice(acgt, d), ref[k:]])]
        region_cigs = ["200M", "%dM%dD%dM" % (k, d, 200 - k - d), "%dM%dI%dM" % (k, d, 200 - k)]
        R = 9
        V = -rng.random((3, R)) * 5
        for r in range(R):
            a = int(np.argmax(V[:, r]))
            V[a, r] += 1.0
            h = region_haps[a]
            s = int(rng.integers(0, len(h) - 70))
            rd = h[s:s + 70].copy()
            rd[int(rng.integers(0, 70))] = rng.choice(acgt)
            reads.append(rd); rcig.append("70M")
        haps += region_haps; hcig += region_cigs; hstart += [0, 0, 0]
        vals.append(V.ravel()); rstart.append(1000 * (g + 1))
        rrs.append(len(reads)); rhs.append(len(haps))
    small = lk.RealignBatch(rrs, rhs, np.concatenate(vals), haps, hcig, hstart, reads, rcig, rstart)
    tiles = 24  # 24 x 108 reads = 2592 > 2048
    t_rrs, t_rhs = [0], [0]
    for t in range(tiles):
        t_rrs += [t * len(reads) + x for x in rrs[1:]]
        t_rhs += [t * len(haps) + x for x in rhs[1:]]
    big = lk.RealignBatch(t_rrs, t_rhs, np.concatenate(vals * tiles), haps * tiles, hcig * tiles, hstart * tiles, reads * tiles,
                          rcig * tiles, rstart * tiles)
    with lk.Context(0) as ctx:
        sb, si, sc, sp = ctx.realign_reads_batch(small)
        bb, bi, bc, bp = ctx.realign_reads_batch(big)
    n = len(reads)
    assert len(bc) == n * tiles and any(c is not None for c in sc)
    for t in range(tiles):
        assert np.array_equal(bb[t * n:(t + 1) * n], sb) and np.array_equal(bi[t * n:(t + 1) * n], si)
        assert np.array_equal(bp[t * n:(t + 1) * n], sp) and bc[t * n:(t + 1) * n] == sc
