"""Shared Smith-Waterman test inputs (CPU oracle tests, emulator tests, GPU parity tests)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "sw_known_answers.json")))
LONG = json.load(open(os.path.join(GOLDEN, "sw_long_pairs.json")))

NEW_SW_PARAMETERS = [200, -150, -260, -11]          # smith_waterman_aligner.rs:17
STANDARD_NGS = [25, -50, -110, -6]                  # :13
ALIGNMENT_TO_BEST_HAPLOTYPE = [10, -15, -30, -5]    # :25
ORIGINAL_DEFAULT = [3, -1, -4, -3]                  # :12
ALL_PARAMS = [NEW_SW_PARAMETERS, STANDARD_NGS, ALIGNMENT_TO_BEST_HAPLOTYPE, ORIGINAL_DEFAULT]
STRATEGIES = ["SoftClip", "InDel", "LeadingInDel", "Ignore"]


def pack(pairs):
    """pairs: list of (ref, alt) byte strings / arrays -> (ref_off, refs, alt_off, alts)."""
    u8 = lambda x: np.frombuffer(x.encode() if isinstance(x, str) else bytes(x), dtype=np.uint8) if not isinstance(x, np.ndarray) else x
    refs = [u8(r) for r, _ in pairs]
    alts = [u8(a) for _, a in pairs]
    ro = np.zeros(len(pairs) + 1, np.int64)
    ao = np.zeros(len(pairs) + 1, np.int64)
    ro[1:] = np.cumsum([len(x) for x in refs])
    ao[1:] = np.cumsum([len(x) for x in alts])
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
    return ro, cat(refs), ao, cat(alts)


def random_pairs(n_pairs, seed, min_len=1, max_len=90, alphabet=b"ACGT"):
    """Mutated copies (SNPs, indels, shifted windows) so every move type and overhang occurs."""
    rng = np.random.default_rng(seed)
    al = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n_pairs):
        n = int(rng.integers(min_len, max_len + 1))
        ref = rng.choice(al, n)
        mode = rng.integers(0, 5)
        if mode == 0:
            alt = rng.choice(al, int(rng.integers(min_len, max_len + 1)))        # unrelated
        else:
            alt = ref.copy()
            for _ in range(int(rng.integers(0, 4))):
                if len(alt) < 2:
                    break
                k = int(rng.integers(0, len(alt)))
                t = rng.integers(0, 3)
                if t == 0:
                    alt[k] = rng.choice(al)
                elif t == 1:
                    alt = np.concatenate([alt[:k], alt[k + int(rng.integers(1, 6)):]])
                else:
                    alt = np.concatenate([alt[:k], rng.choice(al, int(rng.integers(1, 6))), alt[k:]])
            if mode == 2 and len(alt) > 4:
                alt = alt[int(rng.integers(0, len(alt) // 2)):]                       # overhang left
            if mode == 3 and len(alt) > 4:
                alt = np.concatenate([alt, rng.choice(al, int(rng.integers(1, 8)))])  # overhang right
            if mode == 4 and len(alt) > 8:
                a, b = sorted(rng.integers(0, len(alt), 2))
                alt = alt[a:max(b, a + 1)]                                            # exact substring
        if len(alt) == 0:
            alt = rng.choice(al, 1)
        out.append((ref, alt))
    return out


def long_pairs_padded():
    pad = "N" * 10  # CigarUtils::calculate_cigar wraps both in SW_PAD (reads/cigar_utils.rs:10,387-398)
    return [(pad + p["ref"] + pad, pad + p["alt"] + pad) for p in LONG]
