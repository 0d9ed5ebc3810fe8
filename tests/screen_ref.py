"""numpy model of the dense screening pass (csrc/nfc_screen.cu) and of the host-side segment logic.

Test infrastructure: the CUDA kernels are checked against this model, and the model itself is checked against the
exact lane machine (every detector trigger must fall inside a flagged block).
"""
import numpy as np
from scipy.signal import lfilter

BLOCK = 256          # samples per screening block
HALO = 4096          # lane warm-up before its own region (front-end recurrences converge bit-exactly)
PRE_BLOCKS = 2       # active margin before a flagged block
POST_BLOCKS = 4      # active margin after a flagged block
GAP_BLOCKS = 32      # flagged regions closer than this are one segment (GAP_BLOCKS * BLOCK >= 2 * HALO)
START_BLOCKS = 8     # the stream start is always a segment (carrier-off quirk frames, 1024-sample detector gate)


class ScreenParams:
    def __init__(self, rate=10_000_000, corrA=0.75, corrF=0.5, corrV=0.5, modMinB=0.10, power=0.01, margin=0.9):
        stu = rate / float(np.float32(13.56e6))
        self.rate = rate
        self.periods = []
        for r in range(3):
            p1 = int(round(stu * (128 >> r)))
            p2 = int(round(stu * (64 >> r)))
            self.periods.append((p1, p2))
        self.v = (int(round(stu * 256)), int(round(stu * 128)))
        self.kSD = [margin * corrA, margin * min(corrA, corrF), margin * min(corrA, corrF)]
        self.kV = margin * corrV
        # device rule (csrc/nfc_screen.cuh): |C[t] - C[t-q]| > thr * env, thr = min(0.9 T p2, T p2 - 1.25) / 2
        T = [corrA, min(corrA, corrF), min(corrA, corrF)]
        self.thrA = [max(0.25, min(margin * T[r] * self.periods[r][1], T[r] * self.periods[r][1] - 1.25) * 0.5 - (7.5, 2.5, 0.0)[r]) for r in range(3)]
        self.thrV = max(0.25, min(margin * corrV * self.v[1], corrV * self.v[1] - 1.25) - 17.5)
        self.kB = margin * modMinB
        self.low = power / 1.25
        self.high = power * 1.25
        self.envW0 = float(np.float32(1 - 5e5 / rate))
        self.meanW0 = float(np.float32(1 - 5e4 / rate))
        self.level = 0.03


def ema(x, a, init):
    y, _ = lfilter([1 - a], [1, -a], x, zi=[a * init])
    return y


def features(x, sp):
    """per-sample screening features, float64 (the CUDA kernel computes the same quantities in float32)"""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    envf = ema(x, sp.envW0, x[0] if n else 0.0)
    envs = ema(x, sp.meanW0, x[0] if n else 0.0)
    avg = ema(x, sp.meanW0, 0.0)
    w, _ = lfilter([1, -1], [1, -0.9], x, zi=[0.0])  # w[n] = x[n] - x[n-1] + 0.9 w[n-1]; first sample: w = x
    P = np.concatenate([[0.0], np.cumsum(x - (x[0] if n else 0.0))])

    def win(t_shift, p):
        # sum of (x - mu) over [t - t_shift - p + 1, t - t_shift], zero-extended before the stream start
        idx = np.arange(n) + 1 - t_shift
        hi = np.clip(idx, 0, n)
        lo = np.clip(idx - p, 0, n)
        return P[hi] - P[lo]

    sd = []
    dc = []
    for (p1, p2) in sp.periods:
        q = p1 - p2
        c0, cq, c1 = win(0, p2), win(q, p2), win(1, p2)
        sd.append((c0 - 2 * cq + c1) / p2)
        dc.append(c0 - cq)
    p1, p2 = sp.v
    s0v = (win(p1 - p2, p2) - win(0, p2)) / p2
    return dict(x=x, envf=envf, envs=envs, avg=avg, w=w, sd=sd, s0v=s0v, dc=dc, dcv=win(0, p2) - win(p1 - p2, p2))


def sample_flags(x, sp):
    f = features(x, sp)
    env = f["envf"]
    hit = np.zeros(x.size, dtype=bool)
    for r in range(3):
        hit |= np.abs(f["sd"][r]) > sp.kSD[r] * env
    hit |= np.abs(f["s0v"]) > sp.kV * env
    hit |= np.abs(f["w"]) > sp.kB * env
    hit |= np.abs(f["envf"] - f["envs"]) > sp.level * f["envs"]
    hit |= (f["avg"] > 0.9 * sp.low) & (f["avg"] < 1.1 * sp.high)
    return hit, f


def block_flags(x, sp):
    hit, _ = sample_flags(x, sp)
    nb = (x.size + BLOCK - 1) // BLOCK
    pad = np.zeros(nb * BLOCK, dtype=bool)
    pad[:x.size] = hit
    return pad.reshape(nb, BLOCK).any(axis=1)


def active_blocks(raw):
    """dilate raw block flags by the pre / post margins; the stream start is always active"""
    nb = raw.size
    act = np.zeros(nb, dtype=bool)
    idx = np.nonzero(raw)[0]
    for b in idx:
        act[max(0, b - PRE_BLOCKS):min(nb, b + POST_BLOCKS + 1)] = True
    act[:min(nb, START_BLOCKS)] = True
    return act


def segments(act, n):
    """[own_begin, own_end) sample ranges: runs of active blocks, merged across gaps < GAP_BLOCKS"""
    segs = []
    nb = act.size
    b = 0
    while b < nb:
        if not act[b]:
            b += 1
            continue
        e = b
        last = b
        while e < nb and (act[e] or e - last < GAP_BLOCKS):
            if act[e]:
                last = e
            e += 1
        segs.append([b * BLOCK, min(n, (last + 1) * BLOCK)])
        b = last + 1
        while b < nb and not act[b]:
            b += 1
    return segs


def block_flags_device_model(x, sp, band=True):
    """block trigger flags as csrc/nfc_screen.cuh + segment_count_kernel compute them (float64 model):
    per-sample correlator / edge tests against the per-block envelope min(mean(block), mean(previous block)), then the
    block-granular level-shift and carrier-band rules"""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    nb = (n + BLOCK - 1) // BLOCK
    f = features(x, sp)
    xp = np.concatenate([x, np.full(nb * BLOCK - n, x[-1])])
    means = xp.reshape(nb, BLOCK).mean(axis=1)
    prev = np.concatenate([[x[0]], means[:-1]])
    envb = np.maximum(np.minimum(means, prev), 0.0)
    env = np.repeat(envb, BLOCK)[:n]
    hit = np.zeros(n, dtype=bool)
    idx = np.arange(n)
    # decimated evaluation at chunk-relative offsets (17-sample thread chunks of 3840-sample tiles with a 512-sample halo)
    ci = (idx - (idx // 3840) * 3840 + 512) % 17
    hit |= (np.abs(f["dc"][0]) > sp.thrA[0] * env) & ((ci & 3) == 0)
    hit |= (np.abs(f["dc"][1]) > sp.thrA[1] * env) & ((ci & 1) == 0)
    hit |= np.abs(f["dc"][2]) > sp.thrA[2] * env
    hit |= (np.abs(f["dcv"]) > sp.thrV * env) & ((ci & 7) == 0)
    hit |= np.abs(f["w"]) > sp.kB * env
    pad = np.zeros(nb * BLOCK, dtype=bool)
    pad[:n] = hit
    trig = pad.reshape(nb, BLOCK).any(axis=1)
    meanW = sp.meanW0 ** BLOCK
    avg = 0.0
    pm = means[0]
    inband = np.zeros(nb, dtype=bool)
    for b in range(nb):
        m = means[b]
        if abs(m - pm) > 0.025 * max(pm, 1e-6):
            trig[b] = True
        avg_end = meanW * avg + (1 - meanW) * m
        lo, hi = min(avg, avg_end, m), max(avg, avg_end, m)
        if lo < 1.2 * sp.high and hi > 0.8 * sp.low:
            trig[b] = True
            inband[b] = True
        pm, avg = m, avg_end
    if not band:
        return trig
    # one byte per block as the device writes it: bit 0 trigger, bit 3 (SCR_BAND) carrier average near its thresholds
    return trig.astype(np.uint8) | (inband.astype(np.uint8) << 3)
