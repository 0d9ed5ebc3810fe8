#!/usr/bin/env python3
"""Development aid: the host build of the lane pipeline against the compiled reference on a wide corpus (all regression
captures + seeded synthetic streams of every benchmark workload).  Used to vet changes of the segmenting constants.

usage: python tools/cpu_corpus_check.py [streams] [samples] [extra g++ flags for the host build, e.g. -DNFCB200_POST_BLOCKS=2]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import nfcutil as U  # noqa: E402
import screen_ref as S  # noqa: E402
from test_golden_oracle import committed_ref  # noqa: E402
from nfc_laboratory_b200 import synth as Y  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 1_200_000
flags = sys.argv[3:]
if flags:
    so = os.path.join(ROOT, "build", "libhostsim_variant.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-msse2", "-mfpmath=sse", "-ffp-contract=off", "-shared", "-fPIC"] + flags +
                          [os.path.join(ROOT, "tests", "native", "host_sim.cpp"), "-o", so])
    U.HOSTSIM_SO = so
    U.build_hostsim = lambda: None

bad = 0
work = total = 0
t0 = time.time()
for name in U.fixture_names():
    mag, rate, _ = U.fixture_wav(name)
    out, st = U.sim_pipeline(mag, S.block_flags_device_model(mag, S.ScreenParams(rate)), rate)
    ok = out == committed_ref(name)[0]
    bad += not ok
    work += st["work"]; total += mag.size
    if not ok:
        print("MISMATCH", name)
for config in ("nfca106", "nfcb106", "nfca424", "mixed"):
    iq = Y.synth_batch(config, streams, samples, seed=4242).numpy()
    for s in range(streams):
        mag = np.empty(samples, np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq[s]).ctypes.data, mag.size, mag.ctypes.data)
        ref = U.ref_decode(mag, 10_000_000)
        out, st = U.sim_pipeline(mag, S.block_flags_device_model(mag, S.ScreenParams(10_000_000)), 10_000_000)
        work += st["work"]; total += mag.size
        if out != ref:
            bad += 1
            print("MISMATCH", config, s, len(out), len(ref))
print("mismatches: %d, lane samples / samples = %.3f, %.0f s" % (bad, work / total, time.time() - t0))
