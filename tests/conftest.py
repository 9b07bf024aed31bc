import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_pairhmm_golden():
    """Rows of tests/golden/pairhmm-testdata.txt parsed exactly as the reference test does
    (tests/vector_pair_hmm_unit_tests.rs:22-49): phred+33, base quals floored at 6."""
    import numpy as np

    rows = []
    with open(os.path.join(GOLDEN, "pairhmm-testdata.txt")) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            t = line.split()
            if len(t) < 7:
                continue
            hap, read = t[0].encode(), t[1].encode()
            dec = lambda s, lo: np.maximum(np.frombuffer(s.encode(), dtype=np.uint8).astype(np.int32) - 33, lo).astype(np.uint8)
            rows.append(dict(hap=hap, read=read, qual=dec(t[2], 6), ins=dec(t[3], 0), dele=dec(t[4], 0),
                             gcp=dec(t[5], 0), expected=float(t[6])))
    return rows


@pytest.fixture(scope="session")
def pairhmm_golden():
    return load_pairhmm_golden()
