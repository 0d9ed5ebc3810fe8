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

    from bench_standin import StandIn

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


@pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")
def test_two_rank_flow_with_a_stand_in_decoder():
    """the N > 1 path of bench.main() (one process per rank, barrier, frame gather to rank 0, max-over-ranks time) with gloo
    in place of NCCL and the stand-in decoder: rank 0 prints ONE line for the whole job"""
    port = 32500 + (os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")]))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_standin.py"), "--gpus", "2", "--streams", "2", "--samples", "600000",
                                       "--steps", "1", "--warmup", "1"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    lines0 = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    lines1 = [l for l in outs[1][0].splitlines() if l.startswith("{")]
    assert len(lines0) == 1 and lines1 == []
    d = json.loads(lines0[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["cpu_baseline"] is None and d["config"]["sharding"].startswith("streams")
    assert d["decode"]["frames_per_step"] > 0 and d["e2e"]["streams"] == 2 and d["value"] > 0
