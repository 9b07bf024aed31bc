"""Data providers re-expressing the reference's PairHMM unit-test families
(tests/pair_hmm_unit_tests.rs) as plain tuples, shared by the oracle tests (CPU)
and the CUDA parity tests (GPU).

Each case: dict(hap, read, qual, ins, dele, gcp : bytes/np.uint8 arrays,
                tristate: bool, expected: float|None, tol: float|None, name)
"""
import math

import numpy as np

CONTEXT = b"ACGTAATGACGATTGCA"            # pair_hmm_unit_tests.rs CONTEXT
LEFT_FLANK = b"GATTTATCATCGAGTCTGC"
RIGHT_FLANK = b"CATGGATCGTTATCAGCTATCTCGAGGGATTCACTTAACAGTTTTA"
MASSIVE_QUAL = 100
BASES = b"ACGT"


def _full(n, v):
    return np.full(n, v, dtype=np.uint8)


def _with_context(bases, left, right):
    return (LEFT_FLANK if left else b"") + CONTEXT + bases + CONTEXT + (RIGHT_FLANK if right else b"")


def _anchored(n_total, n_read, q, do_gop):
    # BasicLikelihoodTestProvider::qual_as_bytes with anchor_indel = true (:133-156)
    a = _full(n_total, MASSIVE_QUAL)
    if do_gop:
        a[len(CONTEXT)] = q
    else:
        a[len(CONTEXT):len(CONTEXT) + n_read] = q
    return a


def basic_likelihood_cases(extensive=True):
    """make_basic_likelihood_tests (:169-304): 22 080 cases when extensive; tol 0.2 abs."""
    base_quals = [10, 20, 30, 40, 50] if extensive else [30]
    indel_quals = [20, 30, 40, 50] if extensive else [40]
    gcps = [8, 10, 20] if extensive else [10]
    sizes = [2, 3, 4, 5, 7, 8, 9, 10, 20, 30, 35] if extensive else [2]
    out = []

    def add(ref, read, bq, iq, gcp, expected_qual, left, right):
        ref_c = _with_context(ref, left, right)
        read_c = _with_context(read, False, False)
        n = len(read_c)
        expected = expected_qual / -10.0 + 0.03 + math.log10(1.0 / len(ref_c))
        out.append(dict(hap=ref_c, read=read_c, qual=_anchored(n, len(read), bq, False),
                        ins=_anchored(n, len(read), iq, True), dele=_anchored(n, len(read), iq, True),
                        gcp=_anchored(n, len(read), gcp, False), tristate=False, expected=expected, tol=0.2,
                        name="basic bq%d iq%d gcp%d %s/%s l%d r%d" % (bq, iq, gcp, ref[:4], read[:4], left, right)))

    for bq in base_quals:
        for iq in indel_quals:
            for gcp in gcps:
                for rb in BASES:
                    for qb in BASES:
                        add(bytes([rb]), bytes([qb]), bq, iq, gcp, 0 if rb == qb else bq, False, False)
                for size in sizes:
                    for b in BASES:
                        expected_qual = iq + (size - 2) * gcp
                        for insertion_p in (True, False):
                            small, big = bytes([b]), bytes([b]) * size
                            ref, read = (small, big) if insertion_p else (big, small)
                            for left, right in ((False, False), (True, False), (False, True), (True, True)):
                                add(ref, read, bq, iq, gcp, expected_qual, left, right)
    return out


def mismatch_sweep_cases():
    """test_mismatch_in_every_position_in_the_read[_with_centred_haplotype] (:329-409); tol 1e-2."""
    out = []
    for hap, lo, hi in ((b"TTCTCTTCTGTTGTGGCTGGTTTTCTCTTCTGTTGTGGCTGGTTTTCTCTTCTGTTGTGGCTGGTT", 2, 2),
                        (b"TTCTCTTCTGTTGTGGCTGGTT", 2, 0)):
        n = len(hap) - lo - hi
        for k in range(n):
            qual = _full(n, 90)
            qual[k] = 20
            read = bytearray(hap[lo:len(hap) - hi])
            read[k] = ord("T") if read[k] == ord("C") else ord("C")
            expected = math.log10((1.0 / len(hap)) * (1 - 10 ** -9.0) ** (n - 1) * 10 ** -2.0)
            out.append(dict(hap=hap, read=bytes(read), qual=qual, ins=_full(n, 80), dele=_full(n, 80), gcp=_full(n, 80),
                            tristate=False, expected=expected, tol=1e-2, name="mismatch k=%d H=%d" % (k, len(hap))))
    return out


def all_matching_cases():
    """make_hmm_provider -> test_all_matching_read (:436-491) and test_multiple_read_matches_in_haplotype."""
    out = []
    for read_size in (1, 2, 5, 10):
        for ref_size in (1, 2, 5, 10):
            if ref_size <= read_size:
                continue
            read, ref = b"A" * read_size, b"A" * ref_size
            expected = math.log10(abs(ref_size - read_size + 1.0) / ref_size * (1 - 10 ** -2.0) ** read_size)
            out.append(dict(hap=ref, read=read, qual=_full(read_size, 20), ins=_full(read_size, 100),
                            dele=_full(read_size, 100), gcp=_full(read_size, 100), tristate=False,
                            expected=expected, tol=1e-3, name="allmatch %d/%d" % (read_size, ref_size)))
            ref2 = b"CC" + b"A" * ref_size + b"GGA"
            out.append(dict(hap=ref2, read=read, qual=_full(read_size, 20), ins=_full(read_size, 37),
                            dele=_full(read_size, 37), gcp=_full(read_size, 10), tristate=False,
                            expected=None, tol=None, name="multimatch %d/%d" % (read_size, ref_size)))
    for read_size in (1, 2, 5, 10):  # hmm_provider_simple
        read = b"A" * read_size
        out.append(dict(hap=read, read=read, qual=_full(read_size, 20), ins=_full(read_size, 37),
                        dele=_full(read_size, 37), gcp=_full(read_size, 10), tristate=False, expected=None, tol=None,
                        name="same %d" % read_size))
    return out


def big_read_cases():
    """make_big_read_hmm_provider (:542-597): up to 800 x 2000, must just be finite and <= 0."""
    out = []
    read_1, ref_1 = b"ACCAAGTAGTCACCGT", b"ACCAAGTAGTCACCGTAACG"
    for n_read in (1, 2, 10, 20, 50):
        for n_ref in (1, 2, 10, 20, 100):
            if n_ref > n_read:
                read, ref = read_1 * n_read, ref_1 * n_ref
                n = len(read)
                out.append(dict(hap=ref, read=read, qual=_full(n, 30), ins=_full(n, 40), dele=_full(n, 40),
                                gcp=_full(n, 10), tristate=False, expected=None, tol=None,
                                name="big %dx%d" % (len(read), len(ref))))
    return out


def haplotype_indexing_cases():
    """make_haplotype_indexing_provider (:725-814): caching == full recompute; here each (hap, read)
    is simply a full-recompute case (the GPU path never caches)."""
    prefix = b"AACCGGTTTTTGGGCCCAAACGTACGTACAGTTGGTCAACATCGATCAGGTTCCGGAGTAC"
    roots = (b"ACGTGTCAAACCGGGTT", b"ACGTGTCACACTGGGTT", b"ACGTGTCACTCCGCGTT")
    reads = (b"ACGTGTCACACTGGATT", roots[0], roots[1], b"ACGTGTCACACTGGATTCGAT", b"CCAGTAACGTGTCACACTGGATTCGAT")
    out = []
    for read in reads:
        for rl in range(10, len(read)):
            r = read[:rl]
            for ps in range(len(prefix), -1, -6):
                for root in roots:
                    hap = prefix[ps:] + root
                    out.append(dict(hap=hap, read=r, qual=_full(rl, 30), ins=_full(rl, 45), dele=_full(rl, 40),
                                    gcp=_full(rl, 10), tristate=False, expected=None, tol=None, name="hapidx"))
    return out


def pack_cases(cases):
    """Pack independent (read, hap) cases as one-region-per-case batch dict (1 read x 1 hap each)."""
    n = len(cases)
    read_off = np.zeros(n + 1, dtype=np.int64)
    hap_off = np.zeros(n + 1, dtype=np.int64)
    for i, c in enumerate(cases):
        read_off[i + 1] = read_off[i] + len(c["read"])
        hap_off[i + 1] = hap_off[i] + len(c["hap"])
    cat = lambda k: np.concatenate([np.frombuffer(bytes(c[k]), dtype=np.uint8) if isinstance(c[k], (bytes, bytearray))
                                    else np.asarray(c[k], dtype=np.uint8) for c in cases]) if n else np.zeros(0, np.uint8)
    idx = np.arange(n + 1, dtype=np.int64)
    return dict(region_read_start=idx, region_hap_start=idx.copy(), read_off=read_off, hap_off=hap_off,
                read_bases=cat("read"), read_qual=cat("qual"), read_ins=cat("ins"), read_del=cat("dele"),
                read_gcp=cat("gcp"), hap_bases=cat("hap"))
