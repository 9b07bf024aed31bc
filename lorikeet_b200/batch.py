"""Packed batch containers (CSR offsets + byte arenas) — the data layout that crosses the C ABI."""
import ctypes as C

import numpy as np

from . import _lib


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    elif isinstance(a, str):
        a = np.frombuffer(a.encode(), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class PairHmmBatch:
    """A batch of active regions: region g owns reads [region_read_start[g], region_read_start[g+1]) and
    haplotypes [region_hap_start[g], region_hap_start[g+1]); read r owns bytes [read_off[r], read_off[r+1]) of
    the four read arenas.  Mirrors struct lk_pairhmm_batch."""

    FIELDS = ("region_read_start", "region_hap_start", "read_off", "read_bases", "read_qual", "read_ins",
              "read_del", "read_gcp", "hap_off", "hap_bases")

    def __init__(self, region_read_start, region_hap_start, read_off, read_bases, read_qual, read_ins, read_del,
                 hap_off, hap_bases, read_gcp=None, gcp_constant=10):
        self.region_read_start = _i64(region_read_start)
        self.region_hap_start = _i64(region_hap_start)
        self.read_off = _i64(read_off)
        self.read_bases = _u8(read_bases)
        self.read_qual = _u8(read_qual)
        self.read_ins = _u8(read_ins)
        self.read_del = _u8(read_del)
        self.read_gcp = None if read_gcp is None else _u8(read_gcp)
        self.gcp_constant = int(gcp_constant)
        self.hap_off = _i64(hap_off)
        self.hap_bases = _u8(hap_bases)
        # optional: several samples per region (set_samples)
        self.region_sample_start = None
        self.sample_read_start = None

    def set_samples(self, region_sample_start, sample_read_start):
        """Several samples per region: region g owns samples [region_sample_start[g], region_sample_start[g+1]), sample s
        owns reads [sample_read_start[s], sample_read_start[s+1]); the samples of a region tile its reads in order."""
        self.region_sample_start = _i64(region_sample_start)
        self.sample_read_start = _i64(sample_read_start)
        return self

    @classmethod
    def from_dict(cls, d, gcp_constant=None):
        gcp = d.get("read_gcp")
        if gcp_constant is None:
            gcp_constant = int(gcp[0]) if gcp is not None and len(gcp) else 10
        return cls(d["region_read_start"], d["region_hap_start"], d["read_off"], d["read_bases"], d["read_qual"],
                   d["read_ins"], d["read_del"], d["hap_off"], d["hap_bases"], read_gcp=gcp, gcp_constant=gcp_constant)

    # dict-style access so the same object can be handed to the test oracle's batch driver
    def __getitem__(self, k):
        if k == "read_gcp" and self.read_gcp is None:
            return np.full(self.read_bases.shape, self.gcp_constant, dtype=np.uint8)
        return getattr(self, k)

    @property
    def n_regions(self):
        return len(self.region_read_start) - 1

    @property
    def n_reads(self):
        return int(self.region_read_start[-1])

    @property
    def n_haps(self):
        return int(self.region_hap_start[-1])

    def pair_counts(self):
        return np.diff(self.region_read_start) * np.diff(self.region_hap_start)

    @property
    def n_pairs(self):
        return int(self.pair_counts().sum())

    @property
    def n_cells(self):
        rl = np.diff(self.read_off)
        hl = np.diff(self.hap_off)
        rsum = np.add.reduceat(rl, self.region_read_start[:-1]) if len(rl) else np.zeros(self.n_regions, np.int64)
        hsum = np.add.reduceat(hl, self.region_hap_start[:-1]) if len(hl) else np.zeros(self.n_regions, np.int64)
        # reduceat misbehaves on empty segments; regions here are non-empty in every generator
        return int((rsum.astype(np.int64) * hsum.astype(np.int64)).sum())

    def input_bytes(self):
        n = self.read_bases.nbytes * 4 + self.hap_bases.nbytes + self.read_off.nbytes + self.hap_off.nbytes
        n += self.region_read_start.nbytes + self.region_hap_start.nbytes
        if self.read_gcp is not None:
            n += self.read_gcp.nbytes
        return int(n)

    def as_struct(self):
        s = _lib.PairHmmBatchStruct()
        s.n_regions = self.n_regions
        p = lambda a, t: a.ctypes.data_as(t)
        s.region_read_start = p(self.region_read_start, _lib.i64p)
        s.region_hap_start = p(self.region_hap_start, _lib.i64p)
        s.read_off = p(self.read_off, _lib.i64p)
        s.read_bases = p(self.read_bases, _lib.u8p)
        s.read_qual = p(self.read_qual, _lib.u8p)
        s.read_ins = p(self.read_ins, _lib.u8p)
        s.read_del = p(self.read_del, _lib.u8p)
        s.read_gcp = p(self.read_gcp, _lib.u8p) if self.read_gcp is not None else C.cast(None, _lib.u8p)
        s.gcp_constant = self.gcp_constant
        s.hap_off = p(self.hap_off, _lib.i64p)
        s.hap_bases = p(self.hap_bases, _lib.u8p)
        if self.region_sample_start is not None:
            s.region_sample_start = p(self.region_sample_start, _lib.i64p)
            s.sample_read_start = p(self.sample_read_start, _lib.i64p)
        return s

    def region_slice(self, g0, g1):
        """Sub-batch of regions [g0, g1) with offsets rebased to 0 (used by the multi-GPU dispatcher)."""
        r0, r1 = int(self.region_read_start[g0]), int(self.region_read_start[g1])
        h0, h1 = int(self.region_hap_start[g0]), int(self.region_hap_start[g1])
        rb0, rb1 = int(self.read_off[r0]), int(self.read_off[r1])
        hb0, hb1 = int(self.hap_off[h0]), int(self.hap_off[h1])
        return PairHmmBatch(self.region_read_start[g0:g1 + 1] - r0, self.region_hap_start[g0:g1 + 1] - h0,
                            self.read_off[r0:r1 + 1] - rb0, self.read_bases[rb0:rb1], self.read_qual[rb0:rb1],
                            self.read_ins[rb0:rb1], self.read_del[rb0:rb1], self.hap_off[h0:h1 + 1] - hb0,
                            self.hap_bases[hb0:hb1],
                            read_gcp=None if self.read_gcp is None else self.read_gcp[rb0:rb1],
                            gcp_constant=self.gcp_constant)._slice_samples(self, g0, g1, r0)

    def _slice_samples(self, parent, g0, g1, r0):
        if parent.region_sample_start is not None:
            s0, s1 = int(parent.region_sample_start[g0]), int(parent.region_sample_start[g1])
            self.set_samples(parent.region_sample_start[g0:g1 + 1] - s0, parent.sample_read_start[s0:s1 + 1] - r0)
        return self
