"""Shared test utilities: WAV fixtures, ctypes bindings of the oracle (test-only) and of the host simulation.

Nothing here is imported by the product package.
"""
import ctypes as C
import json
import lzma
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE, "_ref", "libnfcref.so")
PORT_SO = os.path.join(ORACLE, "libnfcoracle.so")
HOSTSIM_SO = os.path.join(ROOT, "build", "libhostsim.so")


class RefFrame(C.Structure):
    """oracle/ref_wrap.h nfcref_frame"""
    _fields_ = [
        ("tech_type", C.c_uint32), ("frame_type", C.c_uint32), ("frame_flags", C.c_uint32),
        ("frame_phase", C.c_uint32), ("frame_rate", C.c_uint32), ("length", C.c_uint32),
        ("sample_start", C.c_uint64), ("sample_end", C.c_uint64), ("sample_rate", C.c_uint64),
        ("time_start", C.c_double), ("time_end", C.c_double), ("date_time", C.c_double),
        ("data", C.c_uint8 * 512),
    ]


class SimFrame(C.Structure):
    _fields_ = [
        ("tech", C.c_uint32), ("type", C.c_uint32), ("flags", C.c_uint32), ("phase", C.c_uint32),
        ("rate", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32), ("len", C.c_uint32),
        ("data", C.c_uint8 * 512),
    ]


class SimResult(C.Structure):
    _fields_ = [("stop", C.c_uint32), ("dormant", C.c_uint32), ("locked", C.c_uint32), ("reserved", C.c_uint32)]


def frame_tuple(tech, ftype, flags, phase, rate, start, end, data):
    return (int(tech), int(ftype), int(flags), int(phase), int(rate), int(start), int(end), bytes(data))


def read_wav(path):
    """mono / multi channel 16-bit PCM -> float32 scaled like RecordDevice::readScaledSamples (x / 32768.f)"""
    opener = lzma.open if path.endswith(".xz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos = 12
    fmt = None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", raw[pos + 8:pos + 24])
        elif cid == b"data":
            assert fmt is not None and fmt[5] == 16
            pcm = np.frombuffer(raw, dtype="<i2", count=min(size, len(raw) - pos - 8) // 2, offset=pos + 8)
            return (pcm.astype(np.float32) / np.float32(32768.0)), fmt[2], fmt[1]
        pos += 8 + size + (size & 1)
    raise ValueError("no data chunk in " + path)


def fixture_names():
    names = []
    for fn in sorted(os.listdir(GOLDEN)):
        if fn.endswith(".wav.xz"):
            names.append(fn[:-7])
    return names


def fixture_wav(name):
    return read_wav(os.path.join(GOLDEN, name + ".wav.xz"))


def fixture_golden(name):
    """frames pinned by the reference's own regression JSON (Poll / Listen only)"""
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        frames = json.load(f)["frames"]
    out = []
    for e in frames:
        data = bytes(int(b, 16) for b in e["frameData"].split(":")) if e["frameData"] else b""
        out.append(frame_tuple(e["techType"], e["frameType"], e["frameFlags"], e["framePhase"], e["frameRate"],
                               e["sampleStart"], e["sampleEnd"], data))
    return out


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        lib = C.CDLL(REF_SO)
        lib.nfcref_decode.restype = C.c_long
        lib.nfcref_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint, C.POINTER(RefFrame), C.c_long]
        lib.nfcref_iq_magnitude.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        lib.nfcref_time_batch.restype = C.c_double
        lib.nfcref_time_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_long)]
        _ref = lib
    return _ref


TAP_SO = os.path.join(os.path.dirname(REF_SO), "libnfcref_tap.so")
_tap = None


def tap_lib():
    """the reference decoder with its signal debugger recording float rows in memory (oracle/ref_tap.cpp)"""
    global _tap
    if _tap is None:
        if not os.path.exists(TAP_SO):
            return None
        lib = C.CDLL(TAP_SO)
        lib.nfcref_tap_decode.restype = C.c_long
        lib.nfcref_tap_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint, C.c_void_p, C.c_long]
        _tap = lib
    return _tap


def ref_tap(mag, rate=10000000, enabled=0xF, chunk=65536):
    """[n - 1, 10] float32: the reference's per-sample debug channels (0 x, 1 w, 2 deviation, 3 average, 4.. last writer)"""
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    rows = np.zeros((mag.size, 10), dtype=np.float32)
    n = tap_lib().nfcref_tap_decode(mag.ctypes.data, mag.size, rate, chunk, enabled, rows.ctypes.data, mag.size)
    assert 0 <= n <= mag.size
    return rows[:n]


def sim_trace(mag, rate=10000000, enabled=0xF, first=0, warm=0, stop=None):
    """[stop - first, 8] float32 from the host build of the lane machine: the same channels 0..5 (NaN where this build
    does not tap the value), 6 = lock state after the sample"""
    lib = sim_lib()
    lib.hostsim_trace.restype = C.c_long
    lib.hostsim_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    stop = mag.size if stop is None else min(stop, mag.size)
    rows = np.zeros((stop - first, 8), dtype=np.float32)
    n = lib.hostsim_trace(mag.ctypes.data, stop, rate, enabled, first, warm, rows.ctypes.data)
    assert n == stop - first
    return rows


def ref_decode(mag, rate=10000000, chunk=65536, enabled=0xF, cap=65536):
    """all frames (carrier frames included) of the UNMODIFIED reference decoder"""
    lib = ref_lib()
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    buf = (RefFrame * cap)()
    n = lib.nfcref_decode(mag.ctypes.data, mag.size, rate, chunk, enabled, buf, cap)
    assert n <= cap
    return [frame_tuple(f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start, f.sample_end,
                        bytes(f.data[:f.length])) for f in buf[:n]]


_port = None


def port_lib():
    """the plain-C restatement oracle/nfc_oracle.c (built by oracle/Makefile `port`)"""
    global _port
    if _port is None:
        src = os.path.join(ORACLE, "nfc_oracle.c")
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE, "port"])
        lib = C.CDLL(PORT_SO)
        lib.nfcoracle_decode.restype = C.c_long
        lib.nfcoracle_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint, C.POINTER(RefFrame), C.c_long]
        lib.nfcoracle_iq_magnitude.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _port = lib
    return _port


def port_decode(mag, rate=10000000, enabled=0xF, cap=65536):
    """all frames of the plain-C restatement (same record layout as the compiled reference wrapper)"""
    lib = port_lib()
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    buf = (RefFrame * cap)()
    n = lib.nfcoracle_decode(mag.ctypes.data, mag.size, rate, enabled, buf, cap)
    assert 0 <= n <= cap
    return [frame_tuple(f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start, f.sample_end,
                        bytes(f.data[:f.length])) for f in buf[:n]]


def build_hostsim():
    src = os.path.join(ROOT, "tests", "native", "host_sim.cpp")
    deps = [src] + [os.path.join(ROOT, "nfc_laboratory_b200", "csrc", h) for h in ("nfc_core.h", "nfc_params.h", "nfc_chain.h", "nfc_wlane.h")]
    if os.path.exists(HOSTSIM_SO) and all(os.path.getmtime(HOSTSIM_SO) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(HOSTSIM_SO), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-msse2", "-mfpmath=sse", "-ffp-contract=off", "-shared", "-fPIC",
                           src, "-o", HOSTSIM_SO])


_sim = None


def sim_lib():
    global _sim
    if _sim is None:
        build_hostsim()
        lib = C.CDLL(HOSTSIM_SO)
        lib.hostsim_run.restype = C.c_long
        lib.hostsim_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.POINTER(SimFrame), C.c_long, C.POINTER(SimResult), C.c_void_p, C.c_uint32]
        lib.hostsim_carry_size.restype = C.c_int
        lib.hostsim_pipeline.restype = C.c_long
        lib.hostsim_pipeline.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(SimFrame), C.c_long,
                                         C.POINTER(C.c_uint64), C.c_uint32]
        lib.hostsim_pipeline2.restype = C.c_long
        lib.hostsim_pipeline2.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(SimFrame), C.c_long,
                                          C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32]
        lib.hostsim_set_noff.argtypes = [C.c_int]
        lib.hostsim_set_no_takeover.argtypes = [C.c_int]
        lib.hostsim_set_nofeat.argtypes = [C.c_int]
        lib.hostsim_set_bail.argtypes = [C.c_int]
        lib.hostsim_window.restype = C.c_long
        lib.hostsim_window.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(SimFrame), C.c_long,
                                       C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        lib.hostsim_default_carry.argtypes = [C.c_uint32, C.c_void_p]
        _sim = lib
    return _sim


def sim_run(mag, rate=10000000, enabled=0xF, first=0, warm=4096, own_end=0, carry_in=None, cap=65536, flags=None, block=256):
    """host build of the device lane machine; returns (frames, carry_out bytes, SimResult)"""
    lib = sim_lib()
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    buf = (SimFrame * cap)()
    res = SimResult()
    csz = lib.hostsim_carry_size()
    cout = C.create_string_buffer(csz)
    cin = C.create_string_buffer(carry_in, csz) if carry_in is not None else None
    fl = np.ascontiguousarray(flags, dtype=np.uint8) if flags is not None else None
    n = lib.hostsim_run(mag.ctypes.data, mag.size, rate, enabled, first, warm, own_end, cin, cout, buf, cap, C.byref(res),
                        fl.ctypes.data if fl is not None else None, block)
    assert 0 <= n <= cap
    frames = [frame_tuple(f.tech, f.type, f.flags, f.phase, f.rate, f.start, f.end, bytes(f.data[:f.len])) for f in buf[:n]]
    return frames, cout.raw, res


def sim_pipeline(mag, trigger_blocks, rate=10000000, enabled=0xF, cap=65536, group=1, takeover=True, bail=None):
    """segment-speculative pipeline on the host build of the lane machine; trigger_blocks: bool per 256-sample block"""
    lib = sim_lib()
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    flags = np.ascontiguousarray(trigger_blocks, dtype=np.uint8).copy()
    buf = (SimFrame * cap)()
    stats = (C.c_uint64 * 8)()
    lib.hostsim_set_no_takeover(0 if takeover else 1)
    lib.hostsim_set_bail(-1 if bail is None else int(bail))
    n = lib.hostsim_pipeline(mag.ctypes.data, mag.size, rate, enabled, flags.ctypes.data, flags.size, buf, cap, stats, group)
    lib.hostsim_set_no_takeover(0)
    lib.hostsim_set_bail(-1)
    assert 0 <= n <= cap
    frames = [frame_tuple(f.tech, f.type, f.flags, f.phase, f.rate, f.start, f.end, bytes(f.data[:f.len])) for f in buf[:n]]
    st = dict(lanes=stats[0], live=stats[1], rounds=stats[2], runs=stats[3], work=stats[4], active_blocks=stats[5], bails=stats[6])
    return frames, st


def sim_pipeline2(mag, trigger_blocks, rate=10000000, enabled=0xF, cap=65536, group=0, exact_int=False, noff=False, nofeat=False):
    """round-2 pipeline on the host (front pass -> feature pool -> warp lanes, nfc_wlane.h); group=0: one lane per stream"""
    lib = sim_lib()
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    flags = np.ascontiguousarray(trigger_blocks, dtype=np.uint8).copy()
    buf = (SimFrame * cap)()
    stats = (C.c_uint64 * 8)()
    lib.hostsim_set_noff(1 if noff else 0)
    lib.hostsim_set_nofeat(1 if nofeat else 0)
    n = lib.hostsim_pipeline2(mag.ctypes.data, mag.size, rate, enabled, flags.ctypes.data, flags.size, buf, cap, stats, group, 1 if exact_int else 0)
    lib.hostsim_set_noff(0)
    lib.hostsim_set_nofeat(0)
    assert 0 <= n <= cap
    frames = [frame_tuple(f.tech, f.type, f.flags, f.phase, f.rate, f.start, f.end, bytes(f.data[:f.len])) for f in buf[:n]]
    st = dict(lanes=stats[0], live=stats[1], rounds=stats[2], runs=stats[3], work=stats[4], active_blocks=stats[5], segments=stats[6],
              feature_samples=stats[7])
    return frames, st


def describe(fr):
    return "tech=%x type=%x flags=%02x phase=%x rate=%d [%d..%d] %s" % (fr[0], fr[1], fr[2], fr[3], fr[4], fr[5], fr[6], fr[7].hex(":"))


class HostWindowDecoder:
    """the carry interface of nfc_laboratory_b200.NfcDecoder (set_carry / carry_before / default_carry / carry_size) on the host
    build of the lane pipeline: the per-shard decoder of dist.decode_long_capture_carry in the CPU tests"""

    def __init__(self, rate):
        import screen_ref as S
        self._S = S
        self.rate = rate
        self._carry = None
        self._last = None  # (mag, flags, carry) of the last window
        self.windows = []  # (n_samples, carry injected?) per decode

    def carry_size(self):
        return sim_lib().hostsim_carry_size()

    def default_carry(self):
        buf = C.create_string_buffer(self.carry_size())
        sim_lib().hostsim_default_carry(self.rate, buf)
        return buf.raw

    def set_carry(self, blob, clock_shift=0):
        if blob is None:
            self._carry = None
            return
        b = bytearray(blob)
        assert len(b) == self.carry_size()
        edge = int.from_bytes(b[-4:], "little")
        if edge:
            edge = edge - clock_shift if edge > clock_shift else 1
        b[-4:] = edge.to_bytes(4, "little")
        self._carry = bytes(b)

    def _run(self, query):
        mag, flags, carry = self._last
        lib = sim_lib()
        cap = 65536
        buf = (SimFrame * cap)()
        cout = C.create_string_buffer(self.carry_size())
        begin = C.c_uint32(0)
        fl = flags.copy()
        n = lib.hostsim_window(mag.ctypes.data, mag.size, self.rate, 0xF, fl.ctypes.data, fl.size, buf, cap, carry, min(query, 0xFFFFFFFF), cout,
                               C.byref(begin))
        assert 0 <= n <= cap
        return buf[:n], cout.raw, begin.value

    def decode_batch(self, samples, sigtype=None, sample_rate=None, cap=None):
        from collections import namedtuple
        F = namedtuple("F", "tech_type frame_type frame_flags frame_phase frame_rate sample_start sample_end data")
        mag = np.ascontiguousarray(np.asarray(samples).reshape(-1), dtype=np.float32)
        flags = np.ascontiguousarray(self._S.block_flags(mag, self._S.ScreenParams(self.rate)), dtype=np.uint8)
        self._last = (mag, flags, self._carry)
        self.windows.append((mag.size, self._carry is not None))
        self._carry = None  # one shot
        frames, _, _ = self._run(0)
        return [F(f.tech, f.type, f.flags, f.phase, f.rate, f.start, f.end, bytes(f.data[:f.len])) for f in frames]

    def carry_before(self, sample):
        _, blob, begin = self._run(sample)
        return blob, (None if begin == 0xFFFFFFFF else begin)
