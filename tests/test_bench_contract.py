"""The benchmark's reference arm (`bench.py --impl reference`) runs on host cores only, so its JSON contract is checked here
on a small sample; the GPU arm prints the same keys plus roofline / gpu_launches (driver-side check on the B200)."""
import json
import os
import subprocess
import sys

import pytest

import nfcutil as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--samples", "400000"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "IQ MSamples/s decoded" and d["unit"] == "MSamples/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("nfca106: 1024 synthetic 10 MS/s") and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["frames_per_step"] > 0


def test_host_limits_are_read_from_the_cgroup(tmp_path):
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert 1 <= b.host_cores() <= (os.cpu_count() or 1)
    assert b.host_memory_budget() > 0
    s = b.ClockSampler(0)   # no NVML / nvidia-smi here: the sampler must still start, stop and report
    s.start()
    r = s.summary()
    assert set(r) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_full_size_check_finds_and_explains_a_deviating_stream():
    """bench.py's schedule check, fed with reference-decoded frames standing in for the GPU's: clean batch -> nothing off;
    one frame removed -> that stream is decoded by the reference and reported as different from the 'GPU'"""
    import importlib.util
    import numpy as np
    import torch
    from nfc_laboratory_b200 import synth as Y, dist as ND
    spec = importlib.util.spec_from_file_location("bench_under_test2", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)

    S, n, seed = 3, 900_000, 21
    iq = Y.synth_batch("nfca106", S, n, seed=seed)
    recs = []
    for s in range(S):
        mag = np.empty(n, np.float32)
        U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(iq[s].numpy()).ctypes.data, n, mag.ctypes.data)
        for f in U.ref_decode(mag, 10_000_000):
            recs.append((s,) + tuple(f))
    a = np.zeros(len(recs), dtype=ND.FRAME_DTYPE)
    for i, r in enumerate(recs):
        a[i]["stream"], a[i]["tech_type"], a[i]["frame_type"], a[i]["frame_flags"], a[i]["frame_phase"], a[i]["frame_rate"] = r[:6]
        a[i]["sample_start"], a[i]["sample_end"], a[i]["length"] = r[6], r[7], len(r[8])
        a[i]["data"][: len(r[8])] = np.frombuffer(r[8], dtype=np.uint8)

    res = b.full_size_check(a, S, n, "nfca106", seed, iq)
    assert res["streams_off_schedule"] == 0 and res["expected_poll_listen_frames"] == res["decoded_poll_listen_frames"] > 0

    victim = np.nonzero((a["stream"] == 1) & (a["frame_type"] == 0x103))[0][2]
    res = b.full_size_check(np.delete(a, victim), S, n, "nfca106", seed, iq)
    assert res["streams_off_schedule"] == 1 and res["decoded_poll_listen_frames"] == res["expected_poll_listen_frames"] - 1
    chk = res["deviating_streams_checked"]
    assert len(chk) == 1 and chk[0]["stream"] == 1 and chk[0]["reference"] == chk[0]["expected"] and chk[0]["gpu_equals_reference"] is False
    assert res["gpu_equals_reference_on_them"] is False
    assert b.full_size_check(a, S, n, "mixed", seed, iq) is None


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_gpu_arm_flow_with_a_stand_in_decoder(monkeypatch, capsys):
    """bench.main() end to end on the host: tensors on the CPU and a stand-in for the CUDA decoder that answers with the
    reference's frames.  Checks the control flow and the JSON contract of the GPU arm (resident loop, digest, host-input
    leg, cpu_baseline leg with the oracle spot check and the full-size schedule check) -- not any number in it."""
    import ctypes as C
    import importlib.util
    import numpy as np
    import torch
    import nfc_laboratory_b200 as N
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

    monkeypatch.setenv("NFCB200_BENCH_FLOW_TEST", "1")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a, **k: (1 << 40, 1 << 40))
    monkeypatch.setattr(N, "NfcDecoder", StandIn)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--streams", "3", "--samples", "700000", "--steps", "1", "--warmup", "1"])
    spec = importlib.util.spec_from_file_location("bench_flow", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "phases_ms", "decode", "parity_spot_check", "frames_digest", "full_size_check"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["gpu_launches"] == 9 and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert d["e2e"]["same_frames_as_resident"] is True and d["e2e"]["h2d_bytes_per_step"] == 3 * 700000 * 8 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["parity_spot_check"] is True
    assert d["full_size_check"]["streams_off_schedule"] == 0
    assert d["config"]["workload"].startswith("nfca106: 3 synthetic")
