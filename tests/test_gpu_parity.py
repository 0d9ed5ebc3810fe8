"""Parity of the CUDA path (through the C ABI) against the reference oracle.  Run on the B200 box: pytest -m gpu"""
import numpy as np
import pytest

import nfcutil as U
import screen_ref as S
from test_golden_oracle import committed_ref

pytestmark = pytest.mark.gpu

NAMES = U.fixture_names()


def keys(frames):
    return [f.key() for f in frames]


@pytest.mark.parametrize("name", NAMES)
def test_batch_decode_equals_reference(decoder, name):
    """float magnitude input, one capture per call: every frame field == the reference (carrier frames included)"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    frames = decoder.decode_batch(mag[None, :], N.SIG_MAG_F32, rate)
    assert keys(frames) == committed_ref(name)[0]
    st = decoder.stats()
    assert st["kernel_launches"] >= 5 and st["lanes"] >= 1


@pytest.mark.parametrize("name", NAMES)
def test_poll_listen_frames_equal_golden(decoder, name):
    """what the reference's own regression tool checks (test-sdr main.cpp:171-174, 203-206)"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    frames = decoder.decode_batch(mag[None, :], N.SIG_MAG_F32, rate)
    assert [k for k in keys(frames) if k[1] in (0x102, 0x103)] == U.fixture_golden(name)


@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_001", "test_NFC-B_106kbps_001", "test_NFC-F_212kbps_002", "test_NFC-V_26kbps_002", "test_POLL_ABF_001"])
def test_int16_ingest_equals_reference(decoder, name):
    """WAV int16 decoded on the device (x / 32768.f like RecordDevice.cpp:281-311)"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    pcm = np.round(mag * 32768.0).astype(np.int16)
    assert np.array_equal(pcm.astype(np.float32) / np.float32(32768.0), mag)
    frames = decoder.decode_batch(pcm[None, :], N.SIG_MAG_S16, rate)
    assert keys(frames) == committed_ref(name)[0]


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_002", "test_NFC-A_424kbps_001", "test_NFC-B_106kbps_001", "test_POLL_AB_001"])
def test_iq_input_equals_reference(decoder, name):
    """float2 IQ input: the magnitude step (RadioDeviceTask.cpp:627-637) fused on the device == oracle on sqrtf(I*I+Q*Q)"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    rng = np.random.default_rng(7)
    phi = rng.uniform(0, 2 * np.pi) + np.cumsum(rng.normal(0, 1e-4, mag.size))
    iq = np.stack([mag * np.cos(phi), mag * np.sin(phi)], axis=-1).astype(np.float32)
    refmag = np.empty(mag.size, dtype=np.float32)
    U.ref_lib().nfcref_iq_magnitude(iq.ctypes.data, mag.size, refmag.ctypes.data)
    ref = U.ref_decode(refmag, rate)
    frames = decoder.decode_batch(iq[None], N.SIG_IQ_F32, rate)
    assert keys(frames) == ref


def test_batch_of_streams(decoder):
    """19 captures truncated to a common length as one batch: per-stream frames == reference of the same data"""
    import nfc_laboratory_b200 as N
    if U.ref_lib() is None:
        pytest.skip("oracle not built")
    n = 72944
    data = np.stack([U.fixture_wav(name)[0][:n] for name in NAMES])
    frames = decoder.decode_batch(data, N.SIG_MAG_F32, 10_000_000)
    for s, name in enumerate(NAMES):
        ref = U.ref_decode(data[s], 10_000_000)
        assert [f.key() for f in frames if f.stream == s] == ref, name
    assert [f.stream for f in frames] == sorted(f.stream for f in frames)


@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_001", "test_NFC-A_106kbps_212kbps_001", "test_NFC-B_106kbps_002", "test_NFC-F_212kbps_001",
                                  "test_NFC-V_26kbps_001", "test_POLL_ABF_001"])
def test_streaming_nextframes_equals_reference(name):
    """the reference harness: 65536-sample buffers through nextFrames (test-sdr main.cpp:163-176)"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    d = N.NfcDecoder()
    d.setEnableNfcA(True); d.setEnableNfcB(True); d.setEnableNfcF(True); d.setEnableNfcV(True)
    got = []
    for pos in range(0, mag.size, 65536):
        got += d.nextFrames(mag[pos:pos + 65536], rate)
    got += d.nextFrames(None)
    d.close()
    ref = committed_ref(name)[0]
    assert [f.key() for f in got[:-1]] == ref
    assert got[-1].frame_type in (0x100, 0x101) and got[-1].sample_start == mag.size - 1


def test_tma_and_plain_staging_agree():
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav("test_NFC-F_212kbps_003")
    a = N.NfcDecoder(use_tma=True)
    b = N.NfcDecoder(use_tma=False)
    fa = a.decode_batch(mag[None], N.SIG_MAG_F32, rate)
    fla = a.block_flags()
    fb = b.decode_batch(mag[None], N.SIG_MAG_F32, rate)
    flb = b.block_flags()
    a.close(); b.close()
    assert keys(fa) == keys(fb)
    assert np.array_equal(fla, flb)


def test_screen_flags_cover_numpy_model(decoder):
    """K1 against the numpy model: same trigger set up to float rounding at the threshold (model is float64)"""
    import nfc_laboratory_b200 as N
    for name in ["test_NFC-A_106kbps_003", "test_NFC-B_106kbps_001", "test_NFC-V_26kbps_002"]:
        mag, rate, _ = U.fixture_wav(name)
        decoder.decode_batch(mag[None], N.SIG_MAG_F32, rate)
        flags = decoder.block_flags()[0]
        model = S.block_flags_device_model(mag, S.ScreenParams(rate), band=False)
        trig = (flags & 1).astype(bool)
        disagree = np.count_nonzero(trig != model)
        assert disagree <= max(2, trig.size // 200), (name, disagree, trig.size)


def test_empty_and_invalid_arguments(decoder):
    import nfc_laboratory_b200 as N
    with pytest.raises(N.NfcB200Error) as e:
        decoder.decode_batch_ptr(0, False, N.SIG_MAG_F32, 1, 100, 10_000_000)
    assert e.value.code == -2
    with pytest.raises(N.NfcB200Error) as e:
        decoder.decode_batch(np.zeros((1, 100), dtype=np.float32), N.SIG_MAG_F32, 1000)  # unsupported sample rate
    assert e.value.code == -5
    # a short, ragged stream still decodes (two carrier-off frames)
    frames = decoder.decode_batch(np.zeros((1, 1001), dtype=np.float32), N.SIG_MAG_F32, 10_000_000)
    assert [(f.frame_type, f.sample_start) for f in frames] == [(0x100, 0), (0x100, 1)]


def test_reference_regression_tool_links_against_the_b200_decoder(tmp_path):
    """the reference's UNMODIFIED test-sdr (main.cpp) linked against the lab::NfcDecoder shim + libnfcb200.so prints PASS
    for all 19 captures (built in the container by nfc_laboratory_b200/shim/Makefile, travels as build/dropin/)"""
    import lzma
    import os
    import shutil
    import subprocess
    exe = os.path.join(U.ROOT, "build", "dropin", "test-sdr-b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in harness not built (needs the reference tree)")
    for name in NAMES:
        with lzma.open(os.path.join(U.GOLDEN, name + ".wav.xz"), "rb") as f, open(tmp_path / (name + ".wav"), "wb") as g:
            g.write(f.read())
        shutil.copyfile(os.path.join(U.GOLDEN, name + ".json"), tmp_path / (name + ".json"))
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=900).stdout
    assert out.count("PASS") == 19 and "FAIL" not in out and "UPDATED" not in out, out


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("config", ["nfca106", "nfcb106", "nfca424", "mixed"])
def test_synthetic_workloads_equal_reference(decoder, config):
    """the benchmark's own streams (BASELINE.json configs 2-5 shapes), float2 IQ, 8 streams x 1.5e6 samples as one batch:
    every frame of every stream == the compiled reference on sqrtf(I*I+Q*Q) of the same data"""
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import synth as Y
    iq = Y.synth_batch(config, 8, 1_500_000, seed=77).numpy()
    frames = decoder.decode_batch(iq, N.SIG_IQ_F32, 10_000_000)
    total = 0
    for s in range(iq.shape[0]):
        mag = np.empty(iq.shape[1], np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq[s]).ctypes.data, mag.size, mag.ctypes.data)
        ref = U.ref_decode(mag, 10_000_000)
        total += sum(1 for f in ref if f[1] in (0x102, 0x103))
        assert [f.key() for f in frames if f.stream == s] == ref, (config, s)
    assert total >= 8 * 10


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_001", "test_NFC-B_106kbps_001", "test_NFC-F_212kbps_002"])
def test_int16_iq_ingest_equals_reference(decoder, name):
    """2-channel int16 PCM (a stereo WAV as RecordDevice reads it: x / 32768.f per channel, RecordDevice.cpp:281-311) decoded
    on the device == the oracle on sqrtf(I*I+Q*Q) of the same converted samples"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    rng = np.random.default_rng(11)
    phi = rng.uniform(0, 2 * np.pi) + np.cumsum(rng.normal(0, 1e-4, mag.size))
    pcm = np.stack([np.round(mag * np.cos(phi) * 32767.0), np.round(mag * np.sin(phi) * 32767.0)], axis=-1).astype(np.int16)
    iq = pcm.astype(np.float32) / np.float32(32768.0)
    refmag = np.empty(mag.size, dtype=np.float32)
    U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq).ctypes.data, mag.size, refmag.ctypes.data)
    ref = U.ref_decode(refmag, rate)
    assert sum(1 for f in ref if f[1] in (0x102, 0x103)) >= 4
    frames = decoder.decode_batch(pcm[None], N.SIG_IQ_S16, rate)
    assert keys(frames) == ref


def test_frames_export_roundtrip_from_the_device(decoder, tmp_path):
    """frames decoded on the device -> TRZ container and the regression tool's JSON (export.py) -> the golden file"""
    import json
    import os
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import export as X
    name = "test_NFC-A_424kbps_002"
    mag, rate, _ = U.fixture_wav(name)
    frames = decoder.decode_batch(mag[None, :], N.SIG_MAG_F32, rate)
    X.write_frames_json(tmp_path / "out.json", frames, rate)
    with open(tmp_path / "out.json") as f, open(os.path.join(U.GOLDEN, name + ".json")) as g:
        assert json.load(f) == json.load(g)
    X.write_trz(tmp_path / "out.trz", frames, rate)
    assert X.read_trz(tmp_path / "out.trz") == keys(frames)


@pytest.fixture(scope="module")
def straggler_decoder():
    """a decoder whose thread lanes give up as soon as they run 1 sample past the length they were queued with, whatever the
    queue holds (NFCB200_STRAGGLER=-1, read at nfcb200_create): every such lane is decoded again by a warp lane"""
    import os
    import nfc_laboratory_b200 as N
    old = os.environ.get("NFCB200_STRAGGLER")
    os.environ["NFCB200_STRAGGLER"] = "-1"
    try:
        d = N.NfcDecoder()
    finally:
        if old is None:
            del os.environ["NFCB200_STRAGGLER"]
        else:
            os.environ["NFCB200_STRAGGLER"] = old
    yield d
    d.close()


@pytest.mark.parametrize("name", NAMES)
def test_straggler_pass_equals_reference(straggler_decoder, name):
    """thread lanes + feature-less warp lanes for the lanes that overran + the carry chain over both == the reference"""
    import nfc_laboratory_b200 as N
    mag, rate, _ = U.fixture_wav(name)
    frames = straggler_decoder.decode_batch(mag[None, :], N.SIG_MAG_F32, rate)
    assert keys(frames) == committed_ref(name)[0]
    pcm = np.round(mag * 32768.0).astype(np.int16)
    frames = straggler_decoder.decode_batch(pcm[None, :], N.SIG_MAG_S16, rate)
    assert keys(frames) == committed_ref(name)[0]


def test_straggler_pass_on_a_float_batch(decoder, straggler_decoder):
    """64 synthetic float2 IQ streams: the batch with its overrunning lanes handed to warp lanes decodes to the same frames
    as the plain thread lanes, and lanes were handed over"""
    import torch
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import synth
    iq = synth.synth_batch("nfca106", 64, 2_000_000, seed=5, device="cuda:0")
    a = keys(decoder.decode_batch(iq, N.SIG_IQ_F32, 10_000_000, cap=1 << 18))
    b = keys(straggler_decoder.decode_batch(iq, N.SIG_IQ_F32, 10_000_000, cap=1 << 18))
    st = straggler_decoder.stats()
    assert st["straggler_lanes"] > 0
    assert a == b and len(a) > 1000
