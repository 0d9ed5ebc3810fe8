"""Per-sample values ON THE CUDA PATH against the reference's own signal debugger (oracle/ref_tap.cpp): the device build of
the lane machine (nvcc -fmad=false, device sqrt / div) run with its value taps on (csrc/nfc_trace.cu, nfcb200_debug_trace).
BASELINE.json asks for per-sample correlation values within 1e-5 relative: on the 16-bit regression captures the device is
bit-identical; after a cold start on float input it stays within 1e-5 (the running sums restart with another rounding
history)."""
import ctypes as C

import numpy as np
import pytest

import nfcutil as U

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(U.tap_lib() is None, reason="oracle/_ref/libnfcref_tap.so not built")]

TOL = 1e-5
A_NAMES = [n for n in U.fixture_names() if "NFC-A" in n]


def device_trace(mag, rate, enabled=0xF, first=0, warm=0, stop=None):
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import binding as B
    lib = B.load_library()
    cfg = B.CConfig()
    lib.nfcb200_config_default(C.byref(cfg))
    cfg.enabled = enabled
    mag = np.ascontiguousarray(mag[: (stop + 1) if stop is not None else mag.size], dtype=np.float32)
    rows = np.empty((mag.size - first, 8), dtype=np.float32)
    lib.nfcb200_debug_trace.argtypes = [C.POINTER(B.CConfig), C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    rc = lib.nfcb200_debug_trace(C.byref(cfg), mag.ctypes.data, N.SIG_MAG_F32, mag.size, rate, first, warm, rows.ctypes.data)
    assert rc == 0
    return rows


def markers(ref5, ours5):
    """the debugger overwrites channel 5 with 0.50 / 0.75 on the sync samples (NfcA.cpp:264, 850)"""
    return np.isin(ref5, (np.float32(0.5), np.float32(0.75))) & (ref5 != ours5)


@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_001", "test_NFC-B_106kbps_001", "test_NFC-F_212kbps_002", "test_NFC-V_26kbps_002"])
def test_front_end_signals_are_bit_identical_on_the_device(name):
    mag, rate, _ = U.fixture_wav(name)
    mag = mag[:400_000]
    ref = U.ref_tap(mag, rate)
    ours = device_trace(mag, rate)[: ref.shape[0]]
    for ch in range(4):
        assert np.array_equal(ref[:, ch], ours[:, ch]), (name, ch)


@pytest.mark.parametrize("name", A_NAMES[:3])
def test_running_sums_and_correlation_values_are_bit_identical_on_the_device(name):
    mag, rate, _ = U.fixture_wav(name)
    mag = mag[:400_000]
    ref = U.ref_tap(mag, rate, enabled=1)
    ours = device_trace(mag, rate, enabled=1)[: ref.shape[0]]
    have = ~np.isnan(ours[:, 4])
    assert have.sum() > 0.3 * have.size
    assert np.array_equal(ref[have, 4], ours[have, 4])
    sel = have & ~markers(ref[:, 5], ours[:, 5])
    assert sel.sum() > 0.99 * have.sum()
    assert np.array_equal(ref[sel, 5], ours[sel, 5])
    assert (ours[have, 6] > 0).any() and (ours[have, 6] == 0).any()


def test_cold_started_device_lane_on_float_input_is_within_tolerance():
    from nfc_laboratory_b200 import synth as Y
    iq = Y.synth_batch("nfca106", 1, 400_000, seed=9).numpy()[0]
    mag = np.empty(iq.shape[0], np.float32)
    U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq).ctypes.data, iq.shape[0], mag.ctypes.data)
    first, warm = 150_016, 4096
    ref = U.ref_tap(mag, 10_000_000, enabled=1)[first:]
    ours = device_trace(mag, 10_000_000, enabled=1, first=first, warm=warm)[: ref.shape[0]]
    assert np.array_equal(ref[:, 0], ours[:, 0])
    assert np.array_equal(ref[4096:, 1], ours[4096:, 1]) and np.array_equal(ref[4096:, 2], ours[4096:, 2])
    have = ~np.isnan(ours[:, 5])
    sel = have & ~markers(ref[:, 5], ours[:, 5])
    sel[:warm] = False
    err = np.abs(ref[sel, 5] - ours[sel, 5])
    assert sel.sum() > 10_000
    assert (err <= TOL * np.maximum(np.abs(ref[sel, 5]), ref[sel, 0])).all()
