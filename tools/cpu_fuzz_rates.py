#!/usr/bin/env python3
"""CPU fuzz at other sample rates (4 / 5 / 8 / 12 MS/s; the reference's fixtures are all 10 MS/s): 16-bit quantised synthetic
streams, 1 / 2 / 4 segments per lane, flags from the device's screen model; thread lanes and multi-lane warp lanes of the
host build against the compiled reference (asked twice: it does not always answer the same).

usage: python tools/cpu_fuzz_rates.py <worker> <n_workers> <seconds>
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np, nfcutil as U, screen_ref as S
from nfc_laboratory_b200 import synth
worker=int(sys.argv[1]); nworkers=int(sys.argv[2]); budget=float(sys.argv[3])
t0=time.time(); n=0; bad=0; it=0
wls=("nfca106","nfca424","mixed","nfcb106")
while time.time()-t0 < budget:
    seed = 1300000 + it*nworkers + worker; it+=1
    rng=np.random.default_rng(seed)
    wl = wls[seed % 4]
    amp = float(rng.uniform(0.05, 0.9)); sig = float(np.exp(rng.uniform(np.log(3e-4), np.log(2e-2))))
    ns = int(rng.integers(1_500_000, 3_500_000))
    fs = int(rng.choice([4_000_000, 5_000_000, 8_000_000, 12_000_000]))
    iq = synth.synth_batch(wl, 1, ns, seed=seed, device="cpu", fs=fs, amplitude=(amp, amp*1.2), sigma=(sig, sig*1.5))[0].numpy()
    mag = np.sqrt(iq[:,0].astype(np.float32)**2 + iq[:,1].astype(np.float32)**2).astype(np.float32)
    mag = np.round(np.clip(mag, 0, 0.9999) * 32768.0).astype(np.int16).astype(np.float32) / np.float32(32768.0)
    trig = S.block_flags_device_model(mag, S.ScreenParams(fs))
    r = U.ref_decode(mag, fs)
    group = int(rng.choice([1,2,4]))
    a, sa = U.sim_pipeline(mag, trig, fs, group=group)
    b, sb = U.sim_pipeline2(mag, trig, fs, group=group, exact_int=True)
    n+=1
    if a != r or b != r:
        r2 = U.ref_decode(mag, fs)
        if a != r2 or b != r2:
            bad+=1
            print("DIFF", fs, wl, seed, amp, sig, ns, group, "thread", a==r2, "wlane", b==r2, len(a), len(b), len(r2), flush=True)
    if n % 10 == 0:
        print("worker", worker, "cases", n, "bad", bad, flush=True)
print("worker", worker, "done cases", n, "bad", bad, flush=True)
