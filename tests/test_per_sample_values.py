"""Per-sample values of the decoder against the reference's own signal debugger (oracle/ref_tap.cpp): the front-end
signals, the detectors' running sums and the correlation values S0 / SD.  BASELINE.json asks for correlation values
within 1e-5 relative; from the exact stream start the lane machine is bit-identical, after a cold start it is bit-identical
once the warm-up is over on quantised captures and within 1e-6 on float input.

The lane machine here is the host build of csrc/nfc_core.h (the code the CUDA lanes run); channels 4 / 5 are tapped for
NFC-A (search: filterIntegrate / period2 and the signed correlatedSD of the 424 kbps detector, NfcA.cpp:259-261; poll
symbols: filterIntegrate / period2 and S0 / period4, NfcA.cpp:845-846)."""
import numpy as np
import pytest

import nfcutil as U

pytestmark = pytest.mark.skipif(U.tap_lib() is None, reason="oracle/_ref/libnfcref_tap.so not built")

NAMES = U.fixture_names()
A_NAMES = [n for n in NAMES if "NFC-A" in n]
TOL = 1e-5  # BASELINE.json north_star: per-sample correlation values within 1e-5 relative


def markers(ref5, ours5):
    """the debugger overwrites channel 5 with 0.50 / 0.75 on the sync samples (NfcA.cpp:264, 850)"""
    return np.isin(ref5, (np.float32(0.5), np.float32(0.75))) & (ref5 != ours5)


@pytest.mark.parametrize("name", NAMES)
def test_front_end_signals_are_bit_identical_from_the_stream_start(name):
    """x, DC-removed edge signal, mean deviation, carrier average: every sample, all techs enabled"""
    mag, rate, _ = U.fixture_wav(name)
    ref = U.ref_tap(mag, rate)
    ours = U.sim_trace(mag, rate)[: ref.shape[0]]
    for ch in range(4):
        assert np.array_equal(ref[:, ch], ours[:, ch]), (name, ch)


@pytest.mark.parametrize("name", A_NAMES)
def test_running_sums_and_correlation_values_are_bit_identical_from_the_stream_start(name):
    mag, rate, _ = U.fixture_wav(name)
    ref = U.ref_tap(mag, rate, enabled=1)
    ours = U.sim_trace(mag, rate, enabled=1)[: ref.shape[0]]
    have = ~np.isnan(ours[:, 4])
    assert have.sum() > 0.3 * have.size
    assert np.array_equal(ref[have, 4], ours[have, 4])                      # running sum / period2
    sel = have & ~markers(ref[:, 5], ours[:, 5])
    assert sel.sum() > 0.99 * have.sum()
    assert np.array_equal(ref[sel, 5], ours[sel, 5])                        # correlatedSD (search) / S0 (poll symbols)
    assert (ours[have, 6] > 0).any() and (ours[have, 6] == 0).any()          # both search mode and locked poll frames


@pytest.mark.parametrize("name,first", [("test_NFC-A_106kbps_212kbps_001", 100_000), ("test_NFC-A_106kbps_424kbps_001", 1_000_000),
                                        ("test_NFC-A_424kbps_002", 300_000)])
@pytest.mark.parametrize("warm", [1536, 4096])
def test_cold_started_lane_converges_to_the_reference(name, first, warm):
    """a lane cold-started in the middle of the capture (DESIGN.md "speculation"): DC filter after ~150 samples, deviation
    after ~1000, carrier average after ~3500 (4.5e-4 off after the short warm-up, which is why that one is only used
    away from the carrier thresholds); the correlation values are identical as soon as the detectors run"""
    mag, rate, _ = U.fixture_wav(name)
    stop = min(mag.size - 1, first + 250_000)
    ref = U.ref_tap(mag[: stop + 1], rate, enabled=1)[first:stop]
    ours = U.sim_trace(mag, rate, enabled=1, first=first, warm=warm, stop=stop)
    assert np.array_equal(ref[:, 0], ours[:, 0])
    assert np.array_equal(ref[200:, 1], ours[200:, 1])
    assert np.array_equal(ref[1536:, 2], ours[1536:, 2])
    rel = np.abs(ref[:, 3] - ours[:, 3]) / np.abs(ref[:, 3])
    assert rel[1536:].max() < 1e-3 and np.array_equal(ref[4096:, 3], ours[4096:, 3])
    have = ~np.isnan(ours[:, 5])
    assert np.nonzero(have)[0][0] == warm - 512                             # detectors open 512 samples before the region
    sel = have & ~markers(ref[:, 5], ours[:, 5])
    sel[:warm] = False                                                      # the correlation rings refill over those 512 samples
    err = np.abs(ref[sel, 5] - ours[sel, 5])
    assert (err <= TOL * np.maximum(np.abs(ref[sel, 5]), ref[sel, 0])).all()
    assert np.array_equal(ref[sel, 5], ours[sel, 5])                        # 16-bit captures: the window sums are exact


def test_cold_started_lane_on_float_input_is_within_tolerance():
    """synthetic float IQ (magnitudes are not multiples of 2^-15: the running sums round): after a cold start the
    correlation values differ from the reference's in the last bits only"""
    from nfc_laboratory_b200 import synth as Y
    iq = Y.synth_batch("nfca106", 1, 600_000, seed=9).numpy()[0]
    mag = np.empty(iq.shape[0], np.float32)
    U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq).ctypes.data, mag.size, mag.ctypes.data)
    first, warm = 200_000, 1536
    ref = U.ref_tap(mag, 10_000_000, enabled=1)[first:]
    ours = U.sim_trace(mag, 10_000_000, enabled=1, first=first, warm=warm, stop=first + ref.shape[0])
    have = ~np.isnan(ours[:, 5])
    sel = have & ~markers(ref[:, 5], ours[:, 5])
    sel[:warm] = False
    assert sel.sum() > 250_000
    scale = np.maximum(np.abs(ref[sel, 5]), ref[sel, 0])    # relative to the value, or to the signal level where the value crosses zero
    err = np.abs(ref[sel, 5] - ours[sel, 5]) / scale
    assert err.max() <= TOL
    assert err.max() < 2e-6
