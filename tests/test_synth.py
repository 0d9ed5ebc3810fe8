"""The synthetic benchmark streams decode in the reference exactly as intended, and the
segment-speculative pipeline reproduces the reference on them."""
import numpy as np
import pytest

import nfcutil as U
import screen_ref as S
from nfc_laboratory_b200 import synth as Y

pytestmark = pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")

FS = 10_000_000


@pytest.mark.parametrize("config", ["nfca106", "nfcb106", "nfca424", "mixed"])
def test_templates_decode_as_intended(config):
    rng = np.random.default_rng(1)
    for m, expected in Y.session_templates(config, FS):
        x = np.concatenate([np.ones(30000, np.float32), m, np.ones(60000, np.float32)]) * np.float32(0.30)
        x = np.abs(x + rng.normal(0, 0.0008, x.size)).astype(np.float32)
        frames = [(f[1], f[7]) for f in U.ref_decode(x, FS) if f[1] in (0x102, 0x103)]
        assert frames, config
        if expected is not None:
            assert frames == expected


def test_checksums_known_answers():
    assert Y.crc_a(bytes([0x50, 0x00])) == bytes([0x57, 0xCD])           # HLTA, SURVEY.md Appendix B
    assert Y.crc_b(bytes([0x05, 0x00, 0x00])) == bytes([0x71, 0xFF])     # REQB
    assert Y.crc_f(bytes([0x06, 0x00, 0xFF, 0xFF, 0x00, 0x00])) == bytes([0x09, 0x21])


@pytest.mark.parametrize("config,group", [("nfca106", 1), ("nfca106", 8), ("nfcb106", 1), ("nfca424", 4), ("mixed", 2)])
def test_pipeline_on_synthetic_streams(config, group):
    iq = Y.synth_batch(config, 2, 600_000, seed=11).numpy()
    for s in range(iq.shape[0]):
        mag = np.empty(iq.shape[1], np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq[s]).ctypes.data, mag.size, mag.ctypes.data)
        ref = U.ref_decode(mag, FS)
        assert sum(1 for f in ref if f[1] in (0x102, 0x103)) >= 4
        out, st = U.sim_pipeline(mag, S.block_flags_device_model(mag, S.ScreenParams(FS)), FS, group=group)
        assert out == ref
        assert st["work"] < 1.2 * mag.size


@pytest.mark.parametrize("config", list(Y.SCHEDULE_FIXES_FRAME_COUNT))
def test_expected_frame_count_is_what_the_reference_decodes(config):
    """the full-size property bench.py checks: the number of poll + listen frames follows from the schedule alone"""
    iq = Y.synth_batch(config, 3, 900_000, seed=21).numpy()
    total = 0
    for s in range(iq.shape[0]):
        mag = np.empty(iq.shape[1], np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq[s]).ctypes.data, mag.size, mag.ctypes.data)
        total += sum(1 for f in U.ref_decode(mag, FS) if f[1] in (0x102, 0x103))
    assert total == Y.expected_frame_count(config, 3, 900_000, seed=21) > 0
    assert Y.expected_frame_count("mixed", 3, 900_000, seed=21) is None and Y.expected_frame_count("nfca424", 3, 900_000, seed=21) is None
