"""ctypes binding of include/nfcb200.h.

``NfcDecoder`` mirrors the method names of the reference's ``lab::NfcDecoder``
(src/nfc-lib/lib-lab/lab-radio/src/main/include/lab/nfc/NfcDecoder.h:33-122) so that tests read like the reference's own
harness (src/nfc-test/test-sdr/src/main/cpp/main.cpp:141-180): construct, setEnableNfcX, nextFrames(buffer).
"""
import ctypes as C
import math
import os

import numpy as np

SIG_IQ_F32 = 1
SIG_MAG_F32 = 2
SIG_MAG_S16 = 3
SIG_IQ_S16 = 4

_HERE = os.path.dirname(os.path.abspath(__file__))


class NfcB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("nfcb200 error %d: %s" % (code, message))
        self.code = code


class CFrame(C.Structure):
    _fields_ = [
        ("stream", C.c_uint32), ("tech_type", C.c_uint32), ("frame_type", C.c_uint32), ("frame_flags", C.c_uint32),
        ("frame_phase", C.c_uint32), ("frame_rate", C.c_uint32), ("length", C.c_uint32), ("reserved", C.c_uint32),
        ("sample_start", C.c_uint64), ("sample_end", C.c_uint64), ("sample_rate", C.c_uint64),
        ("time_start", C.c_double), ("time_end", C.c_double), ("date_time", C.c_double),
        ("data", C.c_uint8 * 512),
    ]


class CConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("enabled", C.c_uint32), ("power_level_threshold", C.c_float),
        ("correlation_threshold", C.c_float * 4), ("modulation_min", C.c_float * 4), ("modulation_max", C.c_float * 4),
        ("stream_time", C.c_uint32), ("use_tma", C.c_uint32), ("max_rounds", C.c_uint32), ("segments_per_lane", C.c_uint32),
        ("exact", C.c_uint32), ("reserved", C.c_uint32 * 3),
    ]


class CStats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("blocks", C.c_uint64), ("active_blocks", C.c_uint64), ("segments", C.c_uint64), ("lanes", C.c_uint64),
        ("live_lanes", C.c_uint64), ("lane_runs", C.c_uint64), ("lane_samples", C.c_uint64), ("rounds", C.c_uint64),
        ("frames", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_h2d", C.c_float), ("ms_screen", C.c_float), ("ms_segment", C.c_float), ("ms_lanes", C.c_float),
        ("ms_gather", C.c_float), ("ms_total", C.c_float), ("ms_wall", C.c_float),
        ("ms_front", C.c_float), ("straggler_lanes", C.c_float), ("feature_samples", C.c_uint64),
    ]


class Frame(tuple):
    """(stream, tech_type, frame_type, frame_flags, frame_phase, frame_rate, sample_start, sample_end, data)"""
    __slots__ = ()

    stream = property(lambda s: s[0])
    tech_type = property(lambda s: s[1])
    frame_type = property(lambda s: s[2])
    frame_flags = property(lambda s: s[3])
    frame_phase = property(lambda s: s[4])
    frame_rate = property(lambda s: s[5])
    sample_start = property(lambda s: s[6])
    sample_end = property(lambda s: s[7])
    data = property(lambda s: s[8])

    def key(self):
        """the fields RawFrame::operator== compares (lab-data RawFrame.cpp:82-98), without the stream index"""
        return tuple(self[1:])


EXPORTS = [
    "nfcb200_config_default", "nfcb200_create", "nfcb200_destroy", "nfcb200_configure", "nfcb200_decode_batch",
    "nfcb200_stream_push", "nfcb200_stream_reset", "nfcb200_get_stats", "nfcb200_get_block_flags", "nfcb200_pack_frames",
    "nfcb200_last_error", "nfcb200_version", "nfcb200_device_frames", "nfcb200_emit_records", "nfcb200_stream_pending", "nfcb200_debug_trace", "nfcb200_carry_size", "nfcb200_set_carry",
    "nfcb200_carry_before", "nfcb200_default_carry",
]


def library_path():
    return os.path.join(_HERE, "libnfcb200.so")


_lib = None


def load_library():
    """load libnfcb200.so (built in-tree by __graft_entry__.build() / csrc/Makefile); fails loudly when it is missing"""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NfcB200Error(-1, "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    lib.nfcb200_last_error.restype = C.c_char_p
    lib.nfcb200_version.restype = C.c_char_p
    lib.nfcb200_config_default.argtypes = [C.POINTER(CConfig)]
    lib.nfcb200_create.argtypes = [C.POINTER(CConfig), C.POINTER(C.c_void_p)]
    lib.nfcb200_destroy.argtypes = [C.c_void_p]
    lib.nfcb200_configure.argtypes = [C.c_void_p, C.POINTER(CConfig)]
    lib.nfcb200_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32,
                                         C.POINTER(CFrame), C.c_uint64, C.POINTER(C.c_uint64)]
    lib.nfcb200_stream_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(CFrame), C.c_uint64,
                                        C.POINTER(C.c_uint64)]
    lib.nfcb200_stream_reset.argtypes = [C.c_void_p]
    lib.nfcb200_get_stats.argtypes = [C.c_void_p, C.POINTER(CStats)]
    lib.nfcb200_get_block_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.nfcb200_pack_frames.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.nfcb200_carry_size.restype = C.c_int
    lib.nfcb200_default_carry.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.nfcb200_set_carry.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint32]
    lib.nfcb200_carry_before.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.nfcb200_device_frames.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.nfcb200_emit_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(CFrame),
                                         C.c_uint64, C.POINTER(C.c_uint64)]
    _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise NfcB200Error(rc, lib.nfcb200_last_error().decode("utf-8", "replace"))


_SIG_DTYPE = {SIG_IQ_F32: (np.float32, 2), SIG_MAG_F32: (np.float32, 1), SIG_MAG_S16: (np.int16, 1), SIG_IQ_S16: (np.int16, 2)}


class NfcDecoder:
    """GPU decoder handle.  Method names follow lab::NfcDecoder; batch decoding is the B200-native addition."""

    def __init__(self, device=0, use_tma=True, segments_per_lane=0, exact=False):
        self._lib = load_library()
        self._cfg = CConfig()
        self._lib.nfcb200_config_default(C.byref(self._cfg))
        self._cfg.device = device
        self._cfg.use_tma = 1 if use_tma else 0
        self._cfg.segments_per_lane = segments_per_lane
        self._cfg.exact = 1 if exact else 0
        self._h = C.c_void_p()
        _check(self._lib, self._lib.nfcb200_create(C.byref(self._cfg), C.byref(self._h)))
        self._rate = 0
        self._frames = None
        self._cap = 0

    # --- lifecycle -------------------------------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.nfcb200_destroy(self._h)
            self._h = C.c_void_p()

    cleanup = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self):
        """NfcDecoder::initialize: forget stream state, re-derive parameters at the next buffer"""
        self._apply()
        _check(self._lib, self._lib.nfcb200_stream_reset(self._h))

    def _apply(self):
        _check(self._lib, self._lib.nfcb200_configure(self._h, C.byref(self._cfg)))

    # --- setters / getters of lab::NfcDecoder ------------------------------------------------------------------------
    def _enable(self, bit, on):
        if on:
            self._cfg.enabled |= bit
        else:
            self._cfg.enabled &= ~bit
        self._apply()

    def setEnableNfcA(self, on): self._enable(1, on)
    def setEnableNfcB(self, on): self._enable(2, on)
    def setEnableNfcF(self, on): self._enable(4, on)
    def setEnableNfcV(self, on): self._enable(8, on)
    def isNfcAEnabled(self): return bool(self._cfg.enabled & 1)
    def isNfcBEnabled(self): return bool(self._cfg.enabled & 2)
    def isNfcFEnabled(self): return bool(self._cfg.enabled & 4)
    def isNfcVEnabled(self): return bool(self._cfg.enabled & 8)

    def setSampleRate(self, rate): self._rate = int(rate)
    def sampleRate(self): return self._rate
    def setStreamTime(self, t): self._cfg.stream_time = int(t); self._apply()
    def streamTime(self): return int(self._cfg.stream_time)
    def setPowerLevelThreshold(self, v): self._cfg.power_level_threshold = float(v); self._apply()
    def powerLevelThreshold(self): return float(self._cfg.power_level_threshold)

    def _set_thr(self, t, corr=None, mn=None, mx=None):
        # NaN leaves a value unchanged, like the reference setters (NfcA.cpp:2027-2045)
        if corr is not None and not math.isnan(corr):
            self._cfg.correlation_threshold[t] = corr
        if mn is not None and not math.isnan(mn):
            self._cfg.modulation_min[t] = mn
        if mx is not None and not math.isnan(mx):
            self._cfg.modulation_max[t] = mx
        self._apply()

    def setCorrelationThresholdNfcA(self, v): self._set_thr(0, corr=v)
    def setCorrelationThresholdNfcB(self, v): self._set_thr(1, corr=v)
    def setCorrelationThresholdNfcF(self, v): self._set_thr(2, corr=v)
    def setCorrelationThresholdNfcV(self, v): self._set_thr(3, corr=v)
    def setModulationThresholdNfcA(self, mn, mx): self._set_thr(0, mn=mn, mx=mx)
    def setModulationThresholdNfcB(self, mn, mx): self._set_thr(1, mn=mn, mx=mx)
    def setModulationThresholdNfcF(self, mn, mx): self._set_thr(2, mn=mn, mx=mx)
    def setModulationThresholdNfcV(self, mn, mx): self._set_thr(3, mn=mn, mx=mx)
    def correlationThresholdNfcA(self): return float(self._cfg.correlation_threshold[0])
    def correlationThresholdNfcB(self): return float(self._cfg.correlation_threshold[1])
    def correlationThresholdNfcF(self): return float(self._cfg.correlation_threshold[2])
    def correlationThresholdNfcV(self): return float(self._cfg.correlation_threshold[3])

    # --- decode ------------------------------------------------------------------------------------------------------
    def _buffer(self, cap):
        if cap > self._cap:
            self._frames = (CFrame * cap)()
            self._cap = cap
        return self._frames

    @staticmethod
    def _convert(buf, n):
        out = []
        for i in range(n):
            f = buf[i]
            out.append(Frame((f.stream, f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate,
                              int(f.sample_start), int(f.sample_end), bytes(f.data[:f.length]))))
        return out

    def decode_batch_ptr(self, ptr, on_device, sigtype, n_streams, n_samples, sample_rate, cap=1 << 16, raw=False):
        """decode [n_streams][n_samples] samples at `ptr` (host or device address)"""
        while True:
            buf = self._buffer(cap)
            n = C.c_uint64(0)
            rc = self._lib.nfcb200_decode_batch(self._h, C.c_void_p(ptr), 1 if on_device else 0, sigtype, n_streams, n_samples,
                                                sample_rate, buf, cap, C.byref(n))
            if rc == -4 and n.value > cap:
                cap = int(n.value) + 16
                continue
            _check(self._lib, rc)
            return (buf, n.value) if raw else self._convert(buf, n.value)

    def set_carry(self, blob, clock_shift=0):
        """carry in front of the next single-stream decode (a time shard continuing a capture); None clears"""
        if blob is None:
            _check(self._lib, self._lib.nfcb200_set_carry(self._h, None, 0, 0))
        else:
            b = bytes(blob)
            _check(self._lib, self._lib.nfcb200_set_carry(self._h, b, len(b), int(clock_shift)))

    def carry_size(self):
        return int(self._lib.nfcb200_carry_size())

    def default_carry(self):
        """the carry a cold-started lane assumes in front of it (power-on protocol state, carrier on)"""
        n = self.carry_size()
        buf = C.create_string_buffer(n)
        _check(self._lib, self._lib.nfcb200_default_carry(self._h, buf, n))
        return buf.raw

    def carry_before(self, sample):
        """(carry blob, lane_begin) of the last single-stream decode: the decoder's carry in front of the first lane that
        begins at or after `sample`; lane_begin is None when no lane begins there"""
        n = self._lib.nfcb200_carry_size()
        buf = C.create_string_buffer(n)
        size, begin = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib, self._lib.nfcb200_carry_before(self._h, int(sample), buf, n, C.byref(size), C.byref(begin)))
        return buf.raw, (None if begin.value == 0xFFFFFFFFFFFFFFFF else int(begin.value))

    def device_frames(self):
        """(records_ptr, n_records, ext_ptr, n_ext_chunks): the frames of the last decode_batch as they sit in device memory,
        ordered by (stream, time) -- 128-byte records + 128-byte payload extension chunks (include/nfcb200.h)"""
        rp, ep = C.c_void_p(), C.c_void_p()
        n, ne = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib, self._lib.nfcb200_device_frames(self._h, C.byref(rp), C.byref(n), C.byref(ep), C.byref(ne)))
        return rp.value or 0, int(n.value), ep.value or 0, int(ne.value)

    def emit_records(self, records_ptr, n_records, ext_ptr, n_ext_chunks, stream_offset, sample_rate, raw=False):
        """gathered device-format records (HOST memory) -> frames, stream index raised by stream_offset"""
        buf = (CFrame * max(1, n_records))()
        n = C.c_uint64(0)
        _check(self._lib, self._lib.nfcb200_emit_records(self._h, C.c_void_p(records_ptr), n_records, C.c_void_p(ext_ptr), n_ext_chunks,
                                                         int(stream_offset), int(sample_rate), buf, max(1, n_records), C.byref(n)))
        return (buf, int(n.value)) if raw else self._convert(buf, int(n.value))

    def decode_batch(self, samples, sigtype, sample_rate, cap=1 << 16):
        """samples: numpy array [n_streams, n_samples(, 2)] or torch CUDA tensor of the same shape"""
        dtype, comps = _SIG_DTYPE[sigtype]
        if isinstance(samples, np.ndarray):
            a = np.ascontiguousarray(samples, dtype=dtype)
            if a.ndim == (1 if comps == 1 else 2):
                a = a[None]
            n_streams, n_samples = a.shape[0], a.shape[1]
            return self.decode_batch_ptr(a.ctypes.data, False, sigtype, n_streams, n_samples, sample_rate, cap)
        # torch tensor
        t = samples.contiguous()
        if t.dim() == (1 if comps == 1 else 2):
            t = t[None]
        n_streams, n_samples = t.shape[0], t.shape[1]
        return self.decode_batch_ptr(t.data_ptr(), t.is_cuda, sigtype, n_streams, n_samples, sample_rate, cap)

    def nextFrames(self, samples, sample_rate=None, sigtype=SIG_MAG_F32, cap=4096):
        """NfcDecoder::nextFrames(SignalBuffer): streaming decode of one capture.  samples=None (an invalid buffer in the
        reference, NfcDecoder.cpp:449-463) flushes."""
        rate = int(sample_rate or self._rate or 0)
        buf = self._buffer(cap)
        n = C.c_uint64(0)
        if samples is None:
            rc = self._lib.nfcb200_stream_push(self._h, None, sigtype, 0, rate, buf, cap, C.byref(n))
        else:
            dtype, comps = _SIG_DTYPE[sigtype]
            a = np.ascontiguousarray(samples, dtype=dtype)
            count = a.size // comps
            rc = self._lib.nfcb200_stream_push(self._h, a.ctypes.data, sigtype, count, rate, buf, cap, C.byref(n))
        _check(self._lib, rc)
        self._rate = rate
        return self._convert(buf, n.value)

    def stats(self):
        s = CStats()
        _check(self._lib, self._lib.nfcb200_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in CStats._fields_}

    def block_flags(self):
        nb = C.c_uint64(0)
        _check(self._lib, self._lib.nfcb200_get_block_flags(self._h, None, 0, C.byref(nb)))
        st = self.stats()
        streams = int(st["blocks"] // max(1, nb.value))
        out = np.zeros((streams, nb.value), dtype=np.uint8)
        _check(self._lib, self._lib.nfcb200_get_block_flags(self._h, out.ctypes.data, out.size, C.byref(nb)))
        return out
