"""GPU probe: per-phase timings of the batch decode on the fixtures and on a small synthetic batch (development aid)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import nfcutil as U
import nfc_laboratory_b200 as N
from test_golden_oracle import committed_ref

exact = os.environ.get("PROBE_EXACT") == "1"
d = N.NfcDecoder(exact=exact)
names = U.fixture_names()
if len(sys.argv) > 1:
    names = names[: int(sys.argv[1])]
for name in names:
    mag, rate, _ = U.fixture_wav(name)
    t = time.perf_counter()
    fr = d.decode_batch(mag[None], N.SIG_MAG_F32, rate)
    dt = time.perf_counter() - t
    st = d.stats()
    ok = [f.key() for f in fr] == committed_ref(name)[0]
    print("%-34s %s n=%8d wall=%7.1fms total=%7.1f screen=%6.2f seg=%6.2f lanes=%8.1f gather=%5.2f | segs=%d lanes=%d live=%d rounds=%d runs=%d lane_samples=%.2fx launches=%d" % (
        name, "OK " if ok else "BAD", mag.size, dt * 1e3, st["ms_total"], st["ms_screen"], st["ms_segment"], st["ms_lanes"], st["ms_gather"],
        st["segments"], st["lanes"], st["live_lanes"], st["rounds"], st["lane_runs"], st["lane_samples"] / mag.size, st["kernel_launches"]), flush=True)
d.close()
