"""Streams found by tools/fuzz_parity.py (random exchange sequences of all technologies, gaps from 300 samples up, level steps,
dropouts) on which the lane pipeline once differed from the reference.  Kept as regressions."""
import os
import sys

import numpy as np
import pytest

import nfcutil as U
import screen_ref as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")


def fuzz_stream(seed, index):
    import fuzz_parity as F
    rng = np.random.default_rng(seed)
    tmpl = F.templates()
    x = None
    for _ in range(index + 1):
        x = F.stream(rng, tmpl, int(rng.integers(300_000, 1_500_000)))
    return x, F.FS


@pytest.mark.parametrize("seed,index", [(2, 37), (2, 49), (2, 106), (2, 130), (3, 130), (4, 48), (4, 93), (4, 101)])
def test_dependency_bookkeeping_survives_the_skip_inside_a_lane(seed, index):
    """a lane that skips an idle stretch inside its own region (> 4096 samples between two exchanges of one segment)
    restarts from its own exact carry; what it had recorded about its use of the INCOMING carry (lastCommand written /
    read, NFC-F residue) used to be wiped by that restart, and the carry chain then composed a wrong carry -- e.g. a REQA
    without answer followed 6 000 samples later by another exchange lost lastCommand = REQA, and a later listen frame got
    the application phase instead of the selection phase"""
    x, fs = fuzz_stream(seed, index)
    out, st = U.sim_pipeline(x, S.block_flags_device_model(x, S.ScreenParams(fs)), fs)
    assert out == U.ref_decode(x, fs)
    assert st["lanes"] >= 5
