"""Synthetic captures for decoder paths no reference fixture pins: NFC-B at 212 kbps, NFC-F listen frames at
212 / 424 kbps, NFC-V 1-of-256.  Test infrastructure; the waveform helpers are those of nfc_laboratory_b200/synth.py."""
import numpy as np

from nfc_laboratory_b200.synth import Wave, crc_b, crc_f, _crc16_refl, _nfcb_chars

FS = 10_000_000


def render(w, t_end, amplitude=0.30, sigma=0.0008, lead=60000, tail=80000, seed=1):
    m = w.render(t_end + 4000.0)
    x = np.concatenate([np.ones(lead, np.float32), m, np.ones(tail, np.float32)]) * np.float32(amplitude)
    rng = np.random.default_rng(seed)
    return np.abs(x + rng.normal(0, sigma, x.size)).astype(np.float32)


def nfcb_poll(w, t0, data, rate, depth=0.12):
    """ISO 14443-B poll frame at 106 kbps << rate: SOF 10.5 + 2.5 ETU, characters with start / stop bits, EOF 10.5 ETU"""
    T = 128 >> rate
    lv = 1.0 - depth
    t = t0
    w.low(t, t + 10.5 * T, lv)
    t += 13 * T
    for b in _nfcb_chars(data + crc_b(data)):
        if not b:
            w.low(t, t + T, lv)
        t += T
    w.low(t, t + 10.5 * T, lv)
    return t + 10.5 * T


def nfcf_frame(w, t0, payload, rate, depth):
    """FeliCa frame (poll and listen use the same coding): 48 zero bits, B24D, LEN, payload, CRC, Manchester, MSB first"""
    H = 64 >> rate
    lv = 1.0 - depth
    body = bytes([len(payload) + 1]) + payload
    t = t0
    for b in bytes(6) + b"\xB2\x4D" + body + crc_f(body):
        for k in range(7, -1, -1):
            if (b >> k) & 1:
                w.low(t + H, t + 2 * H, lv)
            else:
                w.low(t, t + H, lv)
            t += 2 * H
    return t


def nfcv_poll_256(w, t0, data, depth=0.98):
    """ISO 15693 1-of-256: SOF pauses in slots 0 and 7 of 8, one pause per byte in slot 2 v + 1 of 512, EOF in slot 2 of 4"""
    U = 128
    lv = 1.0 - depth
    c = _crc16_refl(data, 0xFFFF) ^ 0xFFFF
    t = t0
    w.low(t, t + U, lv)
    w.low(t + 7 * U, t + 8 * U, lv)
    t += 8 * U
    for b in data + bytes([c & 0xFF, c >> 8]):
        w.low(t + (2 * b + 1) * U, t + (2 * b + 2) * U, lv)
        t += 512 * U
    w.low(t + 2 * U, t + 3 * U, lv)
    return t + 4 * U


REQC = bytes([0x00, 0xFF, 0xFF, 0x00, 0x00])
RESC = bytes(range(1, 18))
INVENTORY = bytes([0x26, 0x01, 0x00])


def captures():
    """name -> (samples, [(tech, type, rate, payload)] the reference decodes, flags all zero)"""
    out = {}
    w = Wave(FS)
    t = nfcb_poll(w, 4000.0, bytes([0x05, 0x00, 0x00]), 1)
    out["nfcb_212_poll"] = (render(w, t), [(0x102, 0x102, 211875, bytes([0x05, 0x00, 0x00]) + crc_b(bytes([0x05, 0x00, 0x00])))])
    for rate, sps in ((1, 211875), (2, 423750)):
        w = Wave(FS)
        t = nfcf_frame(w, 4000.0, REQC, rate, 0.40)
        t = nfcf_frame(w, t + 6000.0, RESC, rate, 0.25)
        body_p, body_l = bytes([len(REQC) + 1]) + REQC, bytes([len(RESC) + 1]) + RESC
        out["nfcf_%d_poll_listen" % (sps // 1000)] = (render(w, t, seed=rate), [(0x103, 0x102, sps, body_p + crc_f(body_p)), (0x103, 0x103, sps, body_l + crc_f(body_l))])
    w = Wave(FS)
    t = nfcv_poll_256(w, 4000.0, INVENTORY)
    c = _crc16_refl(INVENTORY, 0xFFFF) ^ 0xFFFF
    out["nfcv_1of256_poll"] = (render(w, t), [(0x104, 0x102, 1655, INVENTORY + bytes([c & 0xFF, c >> 8]))])
    return out
