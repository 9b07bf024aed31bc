"""lorikeet_b200 — B200-native read<->haplotype likelihood hot path (PairHMM + Smith-Waterman).

The product is the CUDA shared library `liblorikeet_b200.so` behind the C ABI of
`include/lorikeet_b200.h`; this package is its thin Python host mirror (ctypes), named after the
reference's modules (src/pair_hmm, src/smith_waterman) so the parity tests read like the reference's.
"""
from ._lib import LorikeetError  # noqa: F401
from .batch import PairHmmBatch  # noqa: F401
from .context import Context, RealignBatch  # noqa: F401
