"""Decoder paths without a reference fixture (NFC-B 212 kbps, NFC-F listen at 212 / 424 kbps, NFC-V 1-of-256): synthetic
captures the compiled reference decodes as intended; the oracle restatement, the single lane and the segment-speculative
lane pipeline (host build of the device code) reproduce the reference frame for frame."""
import numpy as np
import pytest

import extra_signals as X
import nfcutil as U
import screen_ref as S

pytestmark = pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")

CAPS = X.captures()


@pytest.mark.parametrize("name", sorted(CAPS))
def test_reference_decodes_the_capture_as_intended(name):
    x, expected = CAPS[name]
    got = [(f[0], f[1], f[4], f[7]) for f in U.ref_decode(x, X.FS) if f[1] in (0x102, 0x103)]
    assert got == expected
    assert all(f[2] == 0 for f in U.ref_decode(x, X.FS) if f[1] in (0x102, 0x103))


@pytest.mark.parametrize("name", sorted(CAPS))
def test_oracle_restatement_equals_reference(name):
    x, _ = CAPS[name]
    assert U.port_decode(x, X.FS) == U.ref_decode(x, X.FS)


@pytest.mark.parametrize("name", sorted(CAPS))
def test_lane_machine_equals_reference(name):
    x, _ = CAPS[name]
    ref = U.ref_decode(x, X.FS)
    frames, _, _ = U.sim_run(x, X.FS)
    assert frames == ref
    out, st = U.sim_pipeline(x, S.block_flags_device_model(x, S.ScreenParams(X.FS)), X.FS)
    assert out == ref and st["lanes"] >= 2
