"""CPU checks of the round-2 exact pipeline (csrc/nfc_wlane.h) through its host model (tests/native/host_sim.cpp
hostsim_pipeline2): front pass per segment into a feature pool, one WARP LANE per stream (or per group of segments) reading
32 samples at a time, search-mode and locked NFC-A fast paths, exact carry of the running sums over idle stretches (jump on
exact input, sum walk on float).  The bar is the compiled reference, frame for frame, carrier frames included."""
import numpy as np
import pytest

import extra_signals as X
import nfcutil as U
import screen_ref as S
from test_golden_oracle import committed_ref

NAMES = U.fixture_names()
FS = 10_000_000


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("name", NAMES)
def test_fixtures(name, group):
    """all 19 reference captures; group 0 = one lane for the whole capture, 1 = one lane per segment (carry chain)"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    frames, st = U.sim_pipeline2(mag, trig, rate, group=group, exact_int=True)
    assert frames == committed_ref(name)[0]
    if group == 0:
        assert st["lanes"] == 1 and st["rounds"] == 1


@pytest.mark.parametrize("name", NAMES[::3])
def test_fast_paths_change_nothing(name):
    """the same capture with every sample through the per-sample machine (fast-forward paths off) and as float input (sum
    walk instead of the closed-form jump): identical frames"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    a, _ = U.sim_pipeline2(mag, trig, rate, group=0, exact_int=True, noff=True)
    b, _ = U.sim_pipeline2(mag, trig, rate, group=0, exact_int=False)
    assert a == b == committed_ref(name)[0]


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("workload", ["nfca106", "nfcb106", "nfca424", "mixed"])
def test_float_streams_are_exact_with_one_lane_per_stream(workload):
    """float IQ of the benchmark generator (noise: the running sums round): one lane per stream reproduces the reference"""
    from nfc_laboratory_b200 import synth
    iq = synth.synth_batch(workload, 3, 1_200_000, seed=11, device="cpu").numpy()
    for s in range(iq.shape[0]):
        I, Q = iq[s, :, 0], iq[s, :, 1]
        mag = np.sqrt((I * I + Q * Q).astype(np.float32)).astype(np.float32)
        ref = U.ref_decode(mag, FS)
        out, _ = U.sim_pipeline2(mag, S.block_flags_device_model(mag, S.ScreenParams(FS)), FS, group=0)
        assert out == ref


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("name", sorted(X.captures()))
def test_paths_without_a_reference_fixture(name):
    """NFC-B 212 kbps, NFC-F listen 212 / 424 kbps, NFC-V 1-of-256 (tests/extra_signals.py)"""
    x, _ = X.captures()[name]
    ref = U.ref_decode(x, X.FS)
    for group in (0, 1):
        out, _ = U.sim_pipeline2(x, S.block_flags_device_model(x, S.ScreenParams(X.FS)), X.FS, group=group)
        assert out == ref


@pytest.mark.parametrize("name", NAMES)
def test_warp_lanes_without_a_feature_pool(name):
    """the straggler pass of throughput mode: no front pass has run (SegRec.hasFeat == 0), the warp lane computes the
    front-end recurrences itself -- cold at a range's first sample like the front pass, brought up to the entry point after
    a skipped idle stretch (WLane::enter_scalar_cold).  One lane per segment, and one lane for the whole capture on a cut"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    groups = (1, 0) if name in NAMES[::4] else (1,)
    for group in groups:
        frames, _ = U.sim_pipeline2(mag, trig, rate, group=group, exact_int=True, nofeat=True)
        assert frames == committed_ref(name)[0], group


@pytest.mark.parametrize("name", NAMES)
def test_thread_lane_stragglers_handed_to_warp_lanes(name):
    """lanes_kernel + straggler pass on the host: every thread lane that runs past the length it was queued with gives up
    and is decoded again by a feature-less warp lane; its neighbours stay thread lanes; the carry chain joins them"""
    mag, rate, _ = U.fixture_wav(name)
    trig = S.block_flags(mag, S.ScreenParams(rate))
    frames, st = U.sim_pipeline(mag, trig, rate, bail=0)
    assert frames == committed_ref(name)[0], st


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
@pytest.mark.parametrize("workload", ["nfca106", "mixed"])
def test_straggler_pass_on_float_streams(workload):
    """float input (cold-started sums in both kinds of lane): thread lanes, thread lanes with every overrunning lane handed
    to a warp lane, and feature-less warp lanes alone all equal the reference on a synthetic stream"""
    from nfc_laboratory_b200 import synth
    iq = synth.synth_batch(workload, 1, 2_500_000, seed=11, device="cpu")[0].numpy()
    mag = np.sqrt(iq[:, 0].astype(np.float32) ** 2 + iq[:, 1].astype(np.float32) ** 2).astype(np.float32)
    trig = S.block_flags(mag, S.ScreenParams(FS))
    ref = U.ref_decode(mag, FS)
    assert U.sim_pipeline(mag, trig, FS)[0] == ref
    mixed, st = U.sim_pipeline(mag, trig, FS, bail=0)
    assert mixed == ref and st["bails"] > 0, st
    assert U.sim_pipeline2(mag, trig, FS, group=1, exact_int=False, nofeat=True)[0] == ref
