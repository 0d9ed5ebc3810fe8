"""The oracle is pinned against the reference's own golden vectors: wav/test_*.json."""
import json
import os

import pytest

import nfcutil as U

NAMES = U.fixture_names()


def committed_ref(name):
    with open(os.path.join(U.GOLDEN, name + ".ref.json")) as f:
        doc = json.load(f)
    return [U.frame_tuple(t, ty, fl, ph, r, s, e, bytes.fromhex(d)) for (t, ty, fl, ph, r, s, e, d) in doc["frames"]], doc


def test_fixture_inventory():
    assert len(NAMES) == 19
    total = sum(committed_ref(n)[1]["samples"] for n in NAMES)
    assert total == 13954142


@pytest.mark.parametrize("name", NAMES)
def test_committed_reference_output_matches_golden(name):
    """frames recorded from the compiled reference (make_golden.py) == the reference's golden JSON, Poll/Listen only"""
    frames, _ = committed_ref(name)
    assert [f for f in frames if f[1] in (0x102, 0x103)] == U.fixture_golden(name)


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("name", NAMES)
def test_compiled_reference_matches_golden(name):
    mag, rate, _ = U.fixture_wav(name)
    frames = U.ref_decode(mag, rate)
    assert [f for f in frames if f[1] in (0x102, 0x103)] == U.fixture_golden(name)
    assert frames == committed_ref(name)[0]


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_reference_chunk_invariance():
    mag, rate, _ = U.fixture_wav("test_NFC-A_106kbps_001")
    a = U.ref_decode(mag, rate, chunk=65536)
    assert a == U.ref_decode(mag, rate, chunk=1000)
    assert a == U.ref_decode(mag, rate, chunk=7)


@pytest.mark.parametrize("name", NAMES)
def test_c_restatement_matches_golden_and_reference(name):
    """oracle/nfc_oracle.c (plain-C restatement) is pinned against the reference's golden vectors and, frame for frame
    (carrier frames included), against the recorded output of the compiled reference"""
    mag, rate, _ = U.fixture_wav(name)
    frames = U.port_decode(mag, rate)
    assert [f for f in frames if f[1] in (0x102, 0x103)] == U.fixture_golden(name)
    assert frames == committed_ref(name)[0]


def test_c_restatement_iq_magnitude_and_masks():
    import numpy as np
    rng = np.random.default_rng(5)
    iq = rng.normal(0, 0.3, (1000, 2)).astype(np.float32)
    mag = np.empty(1000, dtype=np.float32)
    U.port_lib().nfcoracle_iq_magnitude(iq.ctypes.data, 1000, mag.ctypes.data)
    assert np.array_equal(mag, np.sqrt((iq[:, 0] * iq[:, 0] + iq[:, 1] * iq[:, 1]).astype(np.float32)).astype(np.float32))
    x, rate, _ = U.fixture_wav("test_POLL_ABF_001")
    only_b = U.port_decode(x, rate, enabled=0x2)
    assert only_b and all(f[0] in (0x100, 0x102) for f in only_b)
    assert U.port_decode(np.zeros(5000, np.float32)) == [(0x100, 0x100, 0, 0x101, 0, 0, 0, b""), (0x100, 0x100, 0, 0x101, 0, 1, 1, b"")]
