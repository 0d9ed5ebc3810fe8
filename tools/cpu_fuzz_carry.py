#!/usr/bin/env python3
"""CPU fuzz of the long-capture carry exchange (nfc_laboratory_b200/dist.py decode_long_capture_carry) over the host build of
the lane pipeline (tests/nfcutil.py HostWindowDecoder): random synthetic captures, 2-5 shards, serial form or rank protocol,
candidate spacing 2^16 / 2^18 / 2^20 -- the stitched decode must equal the uncut one.

usage: python tools/cpu_fuzz_carry.py <worker> <n_workers> <seconds>
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np, nfcutil as U
from nfc_laboratory_b200 import synth, dist as ND
worker=int(sys.argv[1]); nworkers=int(sys.argv[2]); budget=float(sys.argv[3])
t0=time.time(); n=0; bad=0; it=0
wls=("nfca106","nfca424","mixed","nfcb106")
key = lambda f: (f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start, f.sample_end, f.data)
while time.time()-t0 < budget:
    seed = 500000 + it*nworkers + worker; it+=1
    rng=np.random.default_rng(seed)
    wl = wls[seed % 4]
    ns = int(rng.integers(2_000_000, 5_000_000))
    shards = int(rng.integers(2, 6))
    ranks = bool(rng.integers(0, 2))
    step = int(rng.choice([1<<16, 1<<18, 1<<20]))
    iq = synth.synth_batch(wl, 1, ns, seed=seed, device="cpu")[0].numpy()
    mag = np.sqrt(iq[:,0].astype(np.float32)**2 + iq[:,1].astype(np.float32)**2).astype(np.float32)
    d = U.HostWindowDecoder(10_000_000)
    full = [key(f) for f in d.decode_batch(mag[None])]
    st={}
    got = ND.decode_long_capture_carry(d, lambda b, e: mag[None, b:e], ns, shards, None, 10_000_000, overlap=1<<18, left=8192, stats=st, model_ranks=ranks, step=step)
    n+=1
    if got != full:
        bad+=1
        print("DIFF", wl, seed, ns, shards, ranks, step, st, len(got), len(full), flush=True)
    if n % 5 == 0:
        print("worker", worker, "cases", n, "bad", bad, st, flush=True)
print("worker", worker, "done cases", n, "bad", bad, flush=True)
