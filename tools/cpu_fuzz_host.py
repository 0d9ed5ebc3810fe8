#!/usr/bin/env python3
"""CPU fuzz of the lane pipeline's host build (tests/native/host_sim.cpp: the device's segment / lane / carry-chain code) against
the compiled reference: random synthetic streams (workload, amplitude 0.05-0.9, noise 3e-4..2e-2, 1.5-4 M samples), decoded by
the speculative thread lanes and by the thread lanes with every overrunning lane handed to a feature-less warp lane.

usage: python tools/cpu_fuzz_host.py <worker> <n_workers> <seconds> [s16]      (test infrastructure: needs oracle/_ref)
s16: quantise every stream to 16 bits first (a WAV capture): the sums are exact there, no stream may differ
exact: the one-lane warp pipeline (exact float sums), with and without a feature pool, instead of the thread lanes
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np, nfcutil as U, screen_ref as S
from nfc_laboratory_b200 import synth
worker=int(sys.argv[1]); nworkers=int(sys.argv[2]); budget=float(sys.argv[3])
t0=time.time(); n=0; bad=0
wls=("nfca106","nfca424","mixed","nfcb106")
it=0
while time.time()-t0 < budget:
    seed = 100000 + it*nworkers + worker
    it+=1
    rng=np.random.default_rng(seed)
    wl = wls[seed % 4]
    amp = float(rng.uniform(0.05, 0.9)); sig = float(np.exp(rng.uniform(np.log(3e-4), np.log(2e-2))))
    ns = int(rng.integers(1_500_000, 4_000_000))
    iq = synth.synth_batch(wl, 1, ns, seed=seed, device="cpu", amplitude=(amp, amp*1.2), sigma=(sig, sig*1.5))[0].numpy()
    mag = np.sqrt(iq[:,0].astype(np.float32)**2 + iq[:,1].astype(np.float32)**2).astype(np.float32)
    if len(sys.argv) > 4 and sys.argv[4] == "s16":
        mag = np.round(np.clip(mag, 0, 0.9999) * 32768.0).astype(np.int16).astype(np.float32) / np.float32(32768.0)
    trig = S.block_flags(mag, S.ScreenParams(10_000_000))
    r = U.ref_decode(mag, 10_000_000)
    if len(sys.argv) > 4 and sys.argv[4] == "exact":
        a, sa = U.sim_pipeline2(mag, trig, 10_000_000, group=0, exact_int=False)
        b, sb = U.sim_pipeline2(mag, trig, 10_000_000, group=0, exact_int=False, nofeat=True)
    else:
        a, sa = U.sim_pipeline(mag, trig, 10_000_000)
        b, sb = U.sim_pipeline(mag, trig, 10_000_000, bail=0)
    n+=1
    if a != r or b != r:
        bad+=1
        d=[i for i,(p,q) in enumerate(zip(a,r)) if p!=q][:2]
        print("DIFF", wl, seed, amp, sig, ns, "thread", a==r, "bail", b==r, len(a), len(r), d, flush=True)
    if n % 10 == 0:
        print("worker", worker, "cases", n, "bad", bad, "rounds", sa['rounds'], flush=True)
print("worker", worker, "done cases", n, "bad", bad, flush=True)
