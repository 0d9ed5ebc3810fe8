"""CPU checks of the host-side logic and of the lane machine compiled for the host (tests/native/host_sim.cpp).

The product runs the lane machine only inside CUDA kernels; the host build exists so that the decoder logic, the
segment construction and the carry chain are exercised in the GPU-less container against the reference's outputs.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nfcutil as U
import screen_ref as S
from test_golden_oracle import committed_ref

NAMES = U.fixture_names()


@pytest.mark.parametrize("name", NAMES)
def test_single_lane_equals_reference(name):
    """one lane over the whole capture == the reference decoder, frame for frame, carrier frames included"""
    mag, rate, _ = U.fixture_wav(name)
    frames, _, _ = U.sim_run(mag, rate)
    assert frames == committed_ref(name)[0]


@pytest.mark.parametrize("name", NAMES)
def test_segment_speculation_equals_reference(name):
    """screen -> segments -> independent cold-started lanes -> carry chain fixed point == the sequential reference"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    frames, st = U.sim_pipeline(mag, trig, rate)
    assert frames == committed_ref(name)[0]
    assert st["rounds"] <= st["lanes"] + 1


def test_screen_is_conservative_on_fixtures():
    """every frame of the reference starts inside an active block of the numpy screen model"""
    for name in NAMES:
        mag, rate, _ = U.fixture_wav(name)
        act = S.active_blocks(S.block_flags(mag, S.ScreenParams(rate)))
        for f in committed_ref(name)[0]:
            if f[1] in (0x102,):
                assert act[min(f[5] // S.BLOCK, act.size - 1)], (name, f)


def test_enable_mask():
    mag, rate, _ = U.fixture_wav("test_POLL_ABF_001")
    frames, _, _ = U.sim_run(mag, rate, enabled=0x1)
    assert frames and all(f[0] in (0x100, 0x101) for f in frames)
    if U.ref_lib() is not None:
        assert frames == U.ref_decode(mag, rate, enabled=0x1)


def test_all_zero_input_quirk():
    """all-zero input gives exactly two NfcCarrierOff frames at samples 0 and 1"""
    frames, _, _ = U.sim_run(np.zeros(50000, dtype=np.float32))
    assert [(f[1], f[5]) for f in frames] == [(0x100, 0), (0x100, 1)]


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads without a GPU and exports every entry point include/nfcb200.h declares"""
    import nfc_laboratory_b200 as N
    header = open(os.path.join(U.ROOT, "include", "nfcb200.h")).read()
    declared = set(re.findall(r"\b(nfcb200_[a-z_]+)\s*\(", header))
    assert declared == set(N.binding.EXPORTS)
    lib = C.CDLL(N.library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to decode (there is no CPU path)"""
    import torch
    import nfc_laboratory_b200 as N
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.NfcB200Error) as e:
        N.NfcDecoder()
    assert e.value.code == -1


def test_struct_layouts_match_header():
    import nfc_laboratory_b200.binding as B
    assert C.sizeof(B.CFrame) == 32 + 24 + 24 + 512
    assert C.sizeof(B.CConfig) == 4 + 4 + 4 + 48 + 4 * 3 + 20


def test_lane_warm_up_is_long_only_near_the_carrier_thresholds():
    """a lane starts 1536 samples before its region, 4096 when a block in which the carrier average can come near its
    thresholds (SCR_BAND, bit 3) lies between the long warm-up start and the sample where the average is exact again"""
    import ctypes as C
    lib = U.sim_lib()
    lib.hostsim_first_sample.restype = C.c_uint32
    lib.hostsim_first_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    nb = 200
    bb = 100

    def first(band_blocks, b=bb):
        f = np.zeros(nb, dtype=np.uint8)
        f[list(band_blocks)] = 9
        return lib.hostsim_first_sample(f.ctypes.data, nb, b)

    assert first([]) == bb * 256 - 1536
    assert first([bb - 16]) == bb * 256 - 4096 and first([bb - 17]) == bb * 256 - 1536
    assert first([bb + 10]) == bb * 256 - 4096 and first([bb + 11]) == bb * 256 - 1536
    assert first([], b=16) == 0 and first([], b=17) == 17 * 256 - 1536


@pytest.mark.parametrize("name", ["test_NFC-V_26kbps_001", "test_NFC-A_424kbps_002", "test_NFC-B_106kbps_002", "test_NFC-F_212kbps_001"])
def test_in_run_takeover_equals_the_chain_walk_extension(name):
    """a lane that is still running when it passes its successor's warm-up start takes that lane's region over in the same
    run (lane_iterate / LaneSucc); without it the chain walk extends the region and the lane runs again: same frames, never
    more rounds"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    a, sa = U.sim_pipeline(mag, trig, rate, takeover=False)
    b, sb = U.sim_pipeline(mag, trig, rate, takeover=True)
    assert a == b
    assert sb["rounds"] <= sa["rounds"] and sb["work"] <= sa["work"] and sb["live"] == sa["live"]


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("workload,rounds,runs", [("nfca424", 2, 1.8), ("nfcb106", 2, 1.8), ("mixed", 3, 2.8)])
def test_carry_chain_converges_in_few_rounds_on_sessions_with_sticky_state(workload, rounds, runs):
    """synthetic sessions that set protocol state (PPS to 424 kbps, ATTRIB parameters) or leave NFC-F search residue behind:
    the speculative lanes + carry chain still equal the sequential reference, and the chain needs no more rounds than the
    dependency depth (a regression here is a throughput loss, not an error: every extra round repeats most lanes)"""
    from nfc_laboratory_b200 import synth
    iq = synth.synth_batch(workload, 1, 2_000_000, seed=5, device="cpu")[0].numpy()
    mag = np.sqrt(iq[:, 0].astype(np.float32) ** 2 + iq[:, 1].astype(np.float32) ** 2).astype(np.float32)
    trig = S.block_flags(mag, S.ScreenParams(10_000_000))
    frames, st = U.sim_pipeline(mag, trig, 10_000_000)
    assert frames == U.ref_decode(mag, 10_000_000)
    assert st["rounds"] <= rounds, st
    assert st["runs"] <= runs * st["lanes"], st  # mixed: a stalled NFC-B SOF search (live residue, NfcB.cpp:308-361) comes and goes with the carry


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_carry_chain_does_not_crawl_behind_a_zero_frame_size():
    """an ATTRIB with an RFU frame-size code leaves maxFrameSize = 0 (NfcB.cpp:1235): the reference truncates every later
    NFC-B frame at its first byte (nothing is emitted) and never assigns the protocol status again.  Every speculative run behind it decodes
    its session and assigns that status -- predicting the carry from what those runs assigned made the chain advance one
    lane per round (98 rounds on the nfcb106 batch); the prediction now passes an inert state through"""
    from nfc_laboratory_b200 import synth as SY
    fs = 10_000_000
    reqb = bytes([0x05, 0x00, 0x00])
    atqb = bytes([0x50, 0x11, 0x22, 0x33, 0x44, 0x00, 0x00, 0x00, 0x00, 0x00, 0x81, 0x81])
    parts = [np.ones(40000, np.float32)]
    for k in range(14):
        attrib = bytes([0x1D, 0x11, 0x22, 0x33, 0x44, 0x00, 0x0D if k == 1 else 0x08, 0x01, 0x00])
        for poll, listen in ((reqb, atqb), (attrib, b"\x00")):
            parts += [SY.nfcb_exchange(fs, poll, listen), np.ones(30000, np.float32)]
    rng = np.random.default_rng(3)
    x = np.concatenate(parts) * np.float32(0.3)
    mag = np.abs(x + rng.normal(0, 0.001, x.size)).astype(np.float32)
    ref = U.ref_decode(mag, fs)
    trig = S.block_flags(mag, S.ScreenParams(fs))
    frames, st = U.sim_pipeline(mag, trig, fs)
    assert frames == ref
    assert ref[-1][7][:7] == bytes([0x1D, 0x11, 0x22, 0x33, 0x44, 0x00, 0x0D]), "nothing decodes behind the RFU ATTRIB: the state is inert"
    assert st["rounds"] <= 3, st


@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_001", "test_NFC-A_424kbps_002", "test_NFC-B_106kbps_002", "test_NFC-F_212kbps_003",
                                  "test_NFC-V_26kbps_001", "test_POLL_ABF_001"])
def test_padding_a_capture_with_its_last_sample_adds_no_frame(name):
    """bench.py's wav_set decodes the 19 captures as one batch call, every capture padded to the longest with its own last
    sample (an idle carrier that goes on): the poll / listen frames stay the golden ones"""
    mag, rate, _ = U.fixture_wav(name)
    x = np.round(mag * 32768.0).astype(np.int16)
    padded = np.empty(2367232, np.int16)
    padded[:x.size] = x
    padded[x.size:] = x[-1]
    pm = padded.astype(np.float32) / np.float32(32768.0)
    frames, _ = U.sim_pipeline(pm, S.block_flags_device_model(pm, S.ScreenParams(rate)), rate)
    assert [k for k in frames if k[1] in (0x102, 0x103)] == U.fixture_golden(name)
