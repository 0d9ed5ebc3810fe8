"""Python model of the segment-speculative decode (what the C++ host orchestration + CUDA lanes do), built from the
numpy screen and the host build of the lane machine.  Test infrastructure only."""
import ctypes as C

import numpy as np

import nfcutil as U
import screen_ref as S


def default_carry(rate):
    lib = U.sim_lib()
    buf = C.create_string_buffer(lib.hostsim_carry_size())
    lib.hostsim_default_carry(rate, buf)
    return buf.raw


def decode_segmented(mag, rate=10_000_000, enabled=0xF, stats=None, halo=S.HALO):
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    n = mag.size
    sp = S.ScreenParams(rate)
    raw = S.block_flags(mag, sp)
    act = S.active_blocks(raw)
    segs = S.segments(act, n)
    spec = default_carry(rate)
    lanes = [dict(begin=b, end=e, carry_in=(None if i == 0 else spec), dirty=True, dead=False) for i, (b, e) in enumerate(segs)]
    flags = act.astype(np.uint8)
    rounds = 0
    runs = 0
    work = 0
    while True:
        dirty = [l for l in lanes if l["dirty"] and not l["dead"]]
        if not dirty:
            break
        rounds += 1
        for l in dirty:
            first = max(0, l["begin"] - halo) if l["carry_in"] is not None or l["begin"] > 0 else 0
            if l["begin"] == 0:
                first = 0
            frames, cout, res = U.sim_run(mag, rate, enabled, first=first, warm=halo, own_end=l["end"], carry_in=l["carry_in"],
                                          flags=flags, block=S.BLOCK)
            l.update(frames=frames, carry_out=cout, stop=res.stop, dirty=False, first=first)
            runs += 1
            work += res.stop - first
        # chain the lanes of this stream
        live = [l for l in lanes if not l["dead"]]
        prev = live[0]
        for l in live[1:]:
            if prev["dirty"]:
                break
            if prev["stop"] + halo > l["begin"]:
                # the previous lane was still busy inside this lane's warm-up zone: it owns this region too
                prev["end"] = max(prev["end"], l["end"])
                l["dead"] = True
                if prev["stop"] < prev["end"]:
                    prev["dirty"] = True
                continue
            if l["carry_in"] != prev["carry_out"]:
                l["carry_in"] = prev["carry_out"]
                l["dirty"] = True
            prev = l
    out = []
    for l in lanes:
        if not l["dead"]:
            out.extend(l["frames"])
    if stats is not None:
        stats.update(rounds=rounds, runs=runs, lanes=len(lanes), live=sum(1 for l in lanes if not l["dead"]), work=work,
                     active=float(act.mean()), n=n)
    return out
