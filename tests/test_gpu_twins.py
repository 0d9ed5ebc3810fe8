"""GPU twins of the CPU-only suites: the SAME inputs as tests/test_edge_inputs.py, tests/test_extra_signals.py and
tests/test_fuzz_regressions.py, decoded by the CUDA library through the C ABI (nfcb200_decode_batch) in both lane modes --
thread lanes (throughput path) and warp lanes (config.exact, one lane per stream) -- and compared with the compiled
reference frame for frame.  The CUDA build is different code generation from the host model of those suites (nvcc
-fmad=false, device sqrt / div, shuffles instead of loops): its equivalence is tested here, not asserted."""
import numpy as np
import pytest

import extra_signals as X
import nfcutil as U
import test_edge_inputs as E
from test_fuzz_regressions import fuzz_stream

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")]

FS = 10_000_000


@pytest.fixture(scope="module", params=[False, True], ids=["thread_lanes", "warp_lanes_exact"])
def dec(request):
    import nfc_laboratory_b200 as N
    d = N.NfcDecoder(exact=request.param)
    d.exact = request.param
    yield d
    d.close()


def gpu(dec, x, rate=FS):
    import nfc_laboratory_b200 as N
    return [f.key() for f in dec.decode_batch(np.ascontiguousarray(x, dtype=np.float32)[None], N.SIG_MAG_F32, rate)]


@pytest.mark.parametrize("name", sorted(E.CASES))
def test_degenerate_inputs(dec, name):
    x = E.CASES[name]
    assert gpu(dec, x) == U.ref_decode(x, FS)


@pytest.mark.parametrize("name", sorted(X.captures()))
def test_paths_without_a_reference_fixture(dec, name):
    """NFC-B 212 kbps, NFC-F listen at 212 / 424 kbps, NFC-V 1-of-256 on the B200"""
    x, expected = X.captures()[name]
    got = gpu(dec, x, X.FS)
    assert got == U.ref_decode(x, X.FS)
    assert [(f[0], f[1], f[4], f[7]) for f in got if f[1] in (0x102, 0x103)] == expected


@pytest.mark.parametrize("seed,index", [(2, 37), (2, 49), (2, 106), (2, 130), (3, 130), (4, 48), (4, 93), (4, 101)])
def test_fuzz_regressions(dec, seed, index):
    x, fs = fuzz_stream(seed, index)
    assert gpu(dec, x, fs) == U.ref_decode(x, fs)


def test_float_batch_is_bit_exact_in_exact_mode():
    """float IQ of the benchmark generator, 24 streams x 1.5e6: the warp lanes reproduce the reference on every stream"""
    import torch
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import synth
    iq = synth.synth_batch("nfca106", 24, 1_500_000, seed=5, device="cuda")
    d = N.NfcDecoder(exact=True)
    got = d.decode_batch(iq, N.SIG_IQ_F32, FS, cap=1 << 18)
    d.close()
    host = iq.cpu().numpy()
    for s in range(host.shape[0]):
        mag = np.empty(host.shape[1], dtype=np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(host[s]).ctypes.data, host.shape[1], mag.ctypes.data)
        assert [f.key() for f in got if f.stream == s] == U.ref_decode(mag, FS), "stream %d" % s
