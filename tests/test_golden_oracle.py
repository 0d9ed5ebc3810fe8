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
