"""Static checks on the built product library (no GPU needed): the kernels the hot paths rely on exist for sm_100a with the
register budgets their occupancy is planned on (cuobjdump -res-usage).  A silent change of a launch bound or a spill-heavy
rebuild shows up here before it shows up as a slower bench line."""
import os
import re
import shutil
import subprocess

import pytest

from lorikeet_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def usage():
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(_lib.SO_PATH):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run([CUOBJDUMP, "-res-usage", _lib.SO_PATH], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert res, "no sm_100a kernels found in the library"
    return res


def _find(usage, fragment):
    hits = {k: v for k, v in usage.items() if fragment in k}
    assert hits, fragment
    return hits


def test_packed_smith_waterman_kernels_fit_sixteen_warps_per_sm(usage):
    # sw_duo_kernel<K, SAME_N>: 4 blocks x 4 warps per SM needs <= 128 registers; no local-memory frame worth mentioning
    for k in (8, 12, 16, 17):
        for same in (0, 1):
            hits = _find(usage, "sw_duo_kernelILi%dELb%dE" % (k, same))
            for name, (regs, stack) in hits.items():
                assert regs <= 128 and stack <= 128, (name, regs, stack)


def test_one_pair_smith_waterman_kernels(usage):
    for k in (2, 4, 8, 12, 16, 17):
        for name, (regs, stack) in _find(usage, "sw_kernelILi%dE" % k).items():
            assert regs <= 128 and stack == 0, (name, regs, stack)


def test_pairhmm_kernels(usage):
    # three blocks of 128 threads per SM: <= 168 registers; the instantiation cfg#2 runs on (K = 19, shared-memory prior table)
    # must not spill at all
    for name, (regs, stack) in _find(usage, "pairhmm_fast_kernel").items():
        assert regs <= 168 and stack <= 128, (name, regs, stack)
    for name, (regs, stack) in _find(usage, "pairhmm_fast_kernelILi19ELi2ELi3E").items():
        assert stack == 0, (name, regs, stack)
    for name, (regs, stack) in _find(usage, "pairhmm_band_kernel").items():
        assert regs <= 168 and stack <= 64, (name, regs, stack)
