"""Degenerate inputs: silence, a bare carrier, carrier steps and ramps around the power threshold, very short and ragged
streams, out-of-range samples.  The compiled reference is the bar for the oracle restatement, the single lane and the
segment-speculative lane pipeline (host build of the device code)."""
import numpy as np
import pytest

import nfcutil as U
import screen_ref as S

pytestmark = pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")

FS = 10_000_000


def noise(n, sigma, seed):
    return np.random.default_rng(seed).normal(0, sigma, n)


def cases():
    c = {}
    c["silence"] = np.zeros(50_000, np.float32)
    c["one_sample"] = np.full(1, 0.3, np.float32)
    c["shorter_than_the_ring"] = np.abs(0.3 + noise(1000, 0.001, 1)).astype(np.float32)
    c["ragged_length"] = np.abs(0.3 + noise(70_001, 0.001, 2)).astype(np.float32)
    c["bare_carrier"] = np.abs(0.35 + noise(200_000, 0.002, 3)).astype(np.float32)
    step = np.concatenate([np.zeros(60_000), np.full(90_000, 0.30), np.zeros(60_000), np.full(50_000, 0.2)])
    c["carrier_on_off_on"] = np.abs(step + noise(step.size, 0.0008, 4)).astype(np.float32)
    ramp = np.concatenate([np.linspace(0, 0.05, 80_000), np.linspace(0.05, 0.0, 80_000), np.full(20_000, 0.011), np.full(20_000, 0.009)])
    c["ramps_around_the_power_threshold"] = np.abs(ramp + noise(ramp.size, 0.0002, 5)).astype(np.float32)
    weak = np.abs(0.0105 + noise(150_000, 0.0004, 6)).astype(np.float32)   # envelope hovering at powerLevelThreshold 0.01
    c["envelope_at_the_detector_gate"] = weak
    big = np.abs(2.5 + noise(100_000, 0.01, 7)).astype(np.float32)         # beyond full scale
    big[40_000:40_030] = 0.0
    big[40_094:40_118] = 0.0
    c["beyond_full_scale_with_pauses"] = big
    neg = (0.3 + noise(100_000, 0.001, 8)).astype(np.float32)
    neg[50_000:50_020] = -0.2                                              # a mono float capture may go negative
    c["negative_samples"] = neg
    spikes = np.abs(0.3 + noise(120_000, 0.001, 9)).astype(np.float32)
    spikes[::9973] = 0.0                                                   # isolated one-sample dropouts
    c["isolated_dropouts"] = spikes
    return c


CASES = cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_restatement_equals_reference(name):
    x = CASES[name]
    assert U.port_decode(x, FS) == U.ref_decode(x, FS)


@pytest.mark.parametrize("name", sorted(CASES))
def test_single_lane_equals_reference(name):
    x = CASES[name]
    frames, _, _ = U.sim_run(x, FS)
    assert frames == U.ref_decode(x, FS)


# The reference stamps a carrier frame with the time of the last strong edge of the DC-removed signal, however old
# (NfcTech.cpp:77-92, NfcDecoder.cpp:477-521): lanes carry that time across their boundaries (Carry::edgeTime) and ignore the
# start-up transient of a cold-started DC filter (Front::edgeHold), so a slow fade inside a later lane is stamped like there.
@pytest.mark.parametrize("name", sorted(CASES))
def test_lane_pipeline_equals_reference(name):
    x = CASES[name]
    if x.size < 2:
        pytest.skip("the block model needs two samples")
    out, _ = U.sim_pipeline(x, S.block_flags_device_model(x, S.ScreenParams(FS)), FS)
    assert out == U.ref_decode(x, FS)


@pytest.mark.parametrize("group", [0, 1])
@pytest.mark.parametrize("name", sorted(CASES))
def test_warp_lane_pipeline_equals_reference(name, group):
    """round-2 pipeline (front pass, feature pool, warp lanes): one lane per stream (group 0) and one lane per segment"""
    x = CASES[name]
    if x.size < 2:
        pytest.skip("the block model needs two samples")
    out, _ = U.sim_pipeline2(x, S.block_flags_device_model(x, S.ScreenParams(FS)), FS, group=group)
    assert out == U.ref_decode(x, FS)


def test_the_cases_exercise_the_carrier_detector():
    kinds = set()
    for name, x in CASES.items():
        for f in U.ref_decode(x, FS):
            kinds.add(f[1])
    assert {0x100, 0x101} <= kinds
    assert [(f[1], f[5]) for f in U.ref_decode(CASES["silence"], FS)] == [(0x100, 0), (0x100, 1)]
