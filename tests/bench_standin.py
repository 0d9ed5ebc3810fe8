"""A stand-in for the CUDA decoder that answers with the reference's frames: lets tests run bench.main() on the host
(tests/test_bench_contract.py).  Test infrastructure."""
import ctypes as C
import os
import sys

import numpy as np

import nfcutil as U
from nfc_laboratory_b200 import binding as B


class StandIn:
    def __init__(self, device=0, **kw):
        self._buf = None
        self._st = {}

    def _decode(self, a, rate):
        out = []
        for s in range(a.shape[0]):
            mag = np.empty(a.shape[1], np.float32)
            U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(a[s]).ctypes.data, a.shape[1], mag.ctypes.data)
            out += [(s,) + tuple(f) for f in U.ref_decode(mag, rate)]
        self._st = {"ms_screen": 1.0, "ms_segment": 0.1, "ms_lanes": 5.0, "ms_gather": 0.1, "ms_total": 6.2, "ms_wall": 6.3, "kernel_launches": 9,
                    "segments": 10, "lanes": 10, "rounds": 1, "lane_runs": 10, "lane_samples": a.shape[0] * a.shape[1] // 4, "samples": a.shape[0] * a.shape[1]}
        return out

    def decode_batch_ptr(self, ptr, on_device, sigtype, S, n, rate, cap=1 << 16, raw=False):
        a = np.ctypeslib.as_array((C.c_float * (S * n * 2)).from_address(ptr)).reshape(S, n, 2)
        recs = self._decode(a, rate)
        buf = (B.CFrame * max(cap, len(recs)))()
        for i, r in enumerate(recs):
            f = buf[i]
            f.stream, f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate = r[:6]
            f.sample_start, f.sample_end, f.sample_rate, f.length = r[6], r[7], rate, len(r[8])
            for k, byte in enumerate(r[8]):
                f.data[k] = byte
        self._buf = buf
        return buf, len(recs)

    def decode_batch(self, t, sigtype, rate, cap=1 << 16):
        return [B.Frame(r) for r in self._decode(t.numpy(), rate)]

    def stats(self):
        return dict(self._st)

    def close(self):
        pass



def run_main():
    """entry of the 2-rank flow test: patch CUDA away, install the stand-in, run bench.main() (argv from the caller)"""
    import importlib.util
    import torch
    import nfc_laboratory_b200 as N
    os.environ["NFCB200_BENCH_FLOW_TEST"] = "1"
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.mem_get_info = lambda *a, **k: (1 << 40, 1 << 40)
    N.NfcDecoder = StandIn
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_flow", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.main()


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    run_main()
