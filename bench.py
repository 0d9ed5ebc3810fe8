#!/usr/bin/env python3
"""Benchmark of the NFC IQ demodulation hot path (BASELINE.json metric: IQ MSamples/s decoded).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S] [--samples n] [--workload nfca106]

One step = one pass of the hot path (IQ -> frames) over one batch of synthetic 10 MS/s IQ streams.  At N = 1 the
default workload is BASELINE.json configs[1]: NFC-A 106 kbps, 1024 streams x 1 s (1e7 samples) of float2 IQ, resident in
HBM (81.9 GB) when the timed region starts.  For N > 1 (torchrun) every rank decodes its own batch of the same shape
(weak scaling, streams are independent) and the frames are gathered to rank 0 over NCCL inside the timed region.

Prints ONE JSON line on rank 0 (contract in the task statement): value = whole-job MSamples/s with device-resident
input; e2e = the same through the C ABI with host buffers (H2D inside the timed region); roofline = the screening
kernel against the measured HBM copy bandwidth; cpu_baseline = the reference's own CPU decoder (oracle/_ref) timed on
this box's cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "IQ MSamples/s decoded"
UNIT = "MSamples/s"
RATE = 10_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--samples", type=int, default=10_000_000)
    ap.add_argument("--workload", default="nfca106", choices=["nfca106", "nfcb106", "nfca424", "mixed"])
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle spot check of the cpu_baseline leg")
    ap.add_argument("--no-wav-set", action="store_true", help="skip the regression-set leg (tests/golden, cpu_baseline leg, untimed)")
    ap.add_argument("--no-full-parity", action="store_true", help="skip the full-size differential against the reference (cpu_baseline leg, untimed)")
    ap.add_argument("--parity-budget", type=float, default=120.0, help="seconds the full-size differential may take")
    ap.add_argument("--exact", action="store_true", help="decode with one warp lane per stream (exact float state, slower)")
    ap.add_argument("--quick", action="store_true", help="small batch for smoke runs (64 streams x 2e6)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md): NVML every 20 ms (a timed region of
    a few hundred milliseconds is over before one nvidia-smi process has started); nvidia-smi is the fallback"""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm = []
        self.max_mhz = None
        self.reasons = set()
        self.power = []
        self.stop_flag = threading.Event()
        self.source = "nvml"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            self.source = "nvidia-smi"

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and index < len(ids) and ids[index].strip().isdigit():
            return int(ids[index])
        return index

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
        except Exception:
            pass
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if len(r) > 1 and r[1].replace(".", "").isdigit():
            self.max_mhz = float(r[1])
        for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None, "source": self.source}


def host_cores():
    """threads the reference arm may use: the scheduler affinity of this process, clipped by a cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def host_memory_budget():
    """bytes of host memory this job may still use: MemAvailable, clipped by the cgroup limit when there is one"""
    avail = 32 << 30
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        pass
    try:
        with open("/sys/fs/cgroup/memory.max") as f:
            lim = f.read().strip()
        if lim != "max":
            with open("/sys/fs/cgroup/memory.current") as f:
                cur = int(f.read().strip())
            avail = min(avail, max(0, int(lim) - cur))
    except Exception:
        pass
    return avail


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(iq_host, threads):
    """the reference's own CPU decoder (unmodified sources, oracle/_ref) on iq_host [S, n, 2] float32: one NfcDecoder per
    host thread, IQ -> magnitude (reference scalar formula) included.  Returns (MSamples/s, frames)."""
    import nfcutil as U
    lib = U.ref_lib()
    if lib is None:
        return None, None
    S, n = iq_host.shape[0], iq_host.shape[1]
    frames = C.c_long(0)
    sec = lib.nfcref_time_batch(None, iq_host.ctypes.data, n, S, RATE, 65536, threads, C.byref(frames))
    return S * n / sec / 1e6, int(frames.value)


def full_parity(frames, iq, S, n, threads, budget_s=90.0, chunk_streams=32):
    """full-size differential against the reference: EVERY stream of the batch is decoded by the reference's own CPU
    decoder (oracle/_ref, `threads` host threads, IQ -> magnitude included) and compared frame for frame -- every field
    RawFrame::operator== compares plus the payload, through one 64-bit hash per frame -- with the frames the GPU returned.
    Part of the cpu_baseline leg (untimed).  Stops early when the time budget is spent and says how far it got."""
    try:
        import torch
        import nfcutil as U
        from nfc_laboratory_b200 import dist as ND
        lib = U.ref_lib()
        if lib is None:
            return None
        lib.nfcref_hash_batch.restype = C.c_double
        lib.nfcref_hash_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
        hashes = ND.frame_hashes(frames)
        streams = frames["stream"].astype(np.int64)
        first = np.searchsorted(streams, np.arange(S + 1))   # frames are ordered by (stream, time)
        cap = int(max(64, (first[1:] - first[:-1]).max() * 2))
        try:
            host = torch.empty((chunk_streams, n, 2), dtype=torch.float32, pin_memory=True)
        except Exception:
            host = torch.empty((chunk_streams, n, 2), dtype=torch.float32)
        rh = np.zeros(chunk_streams * cap, dtype=np.uint64)
        rc = np.zeros(chunk_streams, dtype=np.uint32)
        t0 = time.perf_counter()
        res = {"streams": 0, "frames": 0, "reference_frames": 0, "differing": 0, "streams_differing": 0, "first_diff": None, "reference_seconds": 0.0}
        for c0 in range(0, S, chunk_streams):
            c1 = min(S, c0 + chunk_streams)
            host[:c1 - c0].copy_(iq[c0:c1])
            if iq.is_cuda:
                torch.cuda.synchronize()
            res["reference_seconds"] += float(lib.nfcref_hash_batch(None, host.data_ptr(), n, c1 - c0, RATE, 65536, threads, rh.ctypes.data, cap, rc.ctypes.data))
            for s_ in range(c0, c1):
                mine = hashes[first[s_]:first[s_ + 1]]
                k = int(rc[s_ - c0])
                ref = rh[(s_ - c0) * cap:(s_ - c0) * cap + min(k, cap)]
                m = min(mine.size, ref.size)
                bad = int((mine[:m] != ref[:m]).sum()) + abs(int(mine.size) - k)
                res["frames"] += int(mine.size)
                res["reference_frames"] += k
                if bad:
                    res["differing"] += bad
                    res["streams_differing"] += 1
                    if res["first_diff"] is None:
                        idx = int(np.argmax(mine[:m] != ref[:m])) if m and (mine[:m] != ref[:m]).any() else m
                        res["first_diff"] = {"stream": int(s_), "frame_index": idx, "gpu_frames": int(mine.size), "reference_frames": k}
            res["streams"] = c1
            if time.perf_counter() - t0 > budget_s:
                break
        res["complete"] = res["streams"] == S
        res["seconds"] = time.perf_counter() - t0
        return res
    except Exception as e:  # a diagnostic must not cost the run its number
        return {"error": "%s: %s" % (type(e).__name__, e)}


def lanes_traffic(exact, input_bytes):
    """DRAM bytes per launch of the lane kernel: the ncu capture committed under profiles/ (dram read + write per input byte
    of the captured launch) scaled to this launch's input, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "lanes_kernel_traffic.json")) as f:
            t = json.load(f)
        return float(t["exact" if exact else "thread"]["dram_bytes_per_input_byte"]) * input_bytes
    except Exception:
        return None


def wav_set(dec_factory, threads):
    """the reference's own regression set (19 captures, 13 954 142 samples of 16-bit mono at 10 MS/s,
    tests/golden/) -- decoded capture by capture through nfcb200_decode_batch as int16 and as float, through the streaming
    entry point (nfcb200_stream_push, 65 536-sample chunks like test-sdr), and by the reference on ONE host thread
    (oracle/_ref, how test-sdr runs it) and on all granted threads.  Every GPU frame list is compared with the golden files."""
    try:
        import nfcutil as U
        import nfc_laboratory_b200 as N
        names = U.fixture_names()
        caps = [(nm,) + tuple(U.fixture_wav(nm)[:2]) for nm in names]
        total = sum(c[1].size for c in caps)
        res = {"captures": len(caps), "samples": int(total)}

        def golden_ok(frames, nm):
            return [k for k in frames if k[1] in (0x102, 0x103)] == U.fixture_golden(nm)

        for label, exact in (("batch_thread_lanes", False), ("batch_warp_lanes_exact", True)):
            d = dec_factory(exact)
            ok = True
            for fmt in ("s16", "f32"):
                best = None
                for rep in range(2):  # the first pass includes allocations; two keep the default run near four minutes
                    t0 = time.perf_counter()
                    for nm, mag, rate in caps:
                        if fmt == "s16":
                            x = np.round(mag * 32768.0).astype(np.int16)
                            fr = d.decode_batch(x[None], N.SIG_MAG_S16, rate)
                        else:
                            fr = d.decode_batch(mag[None], N.SIG_MAG_F32, rate)
                        if rep == 0:
                            ok = ok and golden_ok([f.key() for f in fr], nm)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                res["%s_%s_msps" % (label, fmt)] = total / best / 1e6
            res["%s_equals_golden" % label] = bool(ok)
            d.close()

        # the 19 captures as ONE batch call: every capture padded to the longest with its own last sample (an idle carrier that
        # goes on; tests/ checks on the host model that the padding adds no poll / listen frame).  Its own try: a failure here
        # must not cost the other numbers of the leg
        try:
            d = dec_factory(False)
            longest = (max(c[1].size for c in caps) + 255) // 256 * 256
            batch = np.empty((len(caps), longest), dtype=np.int16)
            for i, (nm, mag, rate) in enumerate(caps):
                x = np.round(mag * 32768.0).astype(np.int16)
                batch[i, :x.size] = x
                batch[i, x.size:] = x[-1]
            ok = True
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                fr = d.decode_batch(batch, N.SIG_MAG_S16, caps[0][2], cap=1 << 18)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                if rep == 0:
                    per = [[] for _ in caps]
                    for f in fr:
                        per[f.stream].append(f.key())
                    ok = all(golden_ok(per[i], caps[i][0]) for i in range(len(caps)))
            res["batch_one_call_padded_msps"] = total / best / 1e6  # the captures' own samples per second, padding not counted
            res["batch_one_call_padded_samples"] = int(len(caps) * longest)
            res["batch_one_call_equals_golden"] = bool(ok)
            d.close()
        except Exception as e:
            res["batch_one_call_error"] = "%s: %s" % (type(e).__name__, e)

        # streaming entry point, chunk by chunk
        d = dec_factory(False)
        ok = True
        t0 = time.perf_counter()
        slowest = None
        for nm, mag, rate in caps:
            d.initialize()
            t1 = time.perf_counter()
            out = []
            for pos in range(0, mag.size, 65536):
                out += d.nextFrames(mag[pos:pos + 65536], rate)
            out += d.nextFrames(None, rate)
            dt1 = time.perf_counter() - t1
            rate1 = mag.size / dt1 / 1e6
            slowest = rate1 if slowest is None else min(slowest, rate1)
            ok = ok and golden_ok([f.key() for f in out], nm)
        res["stream_push_msps"] = total / (time.perf_counter() - t0) / 1e6
        res["stream_push_slowest_capture_msps"] = slowest
        res["stream_push_equals_golden"] = bool(ok)
        d.close()

        lib = U.ref_lib()
        if lib is not None:
            def ref_seconds(c):
                fr = C.c_long(0)
                return float(lib.nfcref_time_batch(c[1].ctypes.data, None, c[1].size, 1, c[2], 65536, 1, C.byref(fr)))

            lib.nfcref_time_batch.restype = C.c_double
            res["reference_1_thread_msps"] = total / sum(ref_seconds(c) for c in caps) / 1e6
            import concurrent.futures as cf
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(threads) as ex:
                list(ex.map(ref_seconds, caps))
            res["reference_%d_threads_msps" % threads] = total / (time.perf_counter() - t0) / 1e6
        return res
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def workload_name(workload, S, n, bytes_per_step):
    """config.workload: the same text for both arms (the reference arm times a bounded sample of it, named in config.sample)"""
    return "%s: %d synthetic 10 MS/s x %.1f s float2 IQ streams per GPU (BASELINE.json configs[1] shape), input %.1f GB per GPU, " \
           "larger than L2 (no flush needed)" % (workload, S, n / RATE, bytes_per_step / 1e9)


def full_size_check(frames, S, n, workload, seed, iq, max_streams=4):
    """size-independent property of the whole batch: the number of poll + listen frames of every stream follows from the
    generator's schedule.  Streams that deviate (noise can cost the reference itself a frame) are decoded by the reference
    on the host -- part of the cpu_baseline leg, untimed -- and compared frame for frame with the GPU's.  Never fatal:
    the result is reported as it is."""
    try:
        import nfcutil as U
        from nfc_laboratory_b200 import synth
        exp = synth.expected_frames_per_stream(workload, S, n, seed)
        if exp is None:
            return None
        data = (frames["frame_type"] == 0x102) | (frames["frame_type"] == 0x103)
        got = np.bincount(frames["stream"][data].astype(np.int64), minlength=S)[:S]
        off = np.nonzero(got != exp)[0]
        res = {"expected_poll_listen_frames": int(exp.sum()), "decoded_poll_listen_frames": int(got.sum()), "streams_off_schedule": int(off.size)}
        if off.size and U.ref_lib() is not None:
            agree = True
            checked = []
            for s_ in off[:max_streams]:
                one = np.ascontiguousarray(iq[int(s_)].cpu().numpy())
                mag = np.empty(n, dtype=np.float32)
                U.ref_lib().nfcref_iq_magnitude(one.ctypes.data, n, mag.ctypes.data)
                ref = U.ref_decode(mag, RATE)
                rows = frames[frames["stream"] == s_]
                mine = [(int(r["tech_type"]), int(r["frame_type"]), int(r["frame_flags"]), int(r["frame_phase"]), int(r["frame_rate"]),
                         int(r["sample_start"]), int(r["sample_end"]), bytes(r["data"][: int(r["length"])])) for r in rows]
                same = mine == [tuple(f) for f in ref]
                agree = agree and same
                checked.append({"stream": int(s_), "expected": int(exp[s_]), "decoded": int(got[s_]),
                                "reference": sum(1 for f in ref if f[1] in (0x102, 0x103)), "gpu_equals_reference": bool(same)})
            res["deviating_streams_checked"] = checked
            res["gpu_equals_reference_on_them"] = bool(agree)
        return res
    except Exception as e:  # a diagnostic must not cost the run its number
        return {"error": "%s: %s" % (type(e).__name__, e)}


def reference_arm(args, rank, world):
    """--impl reference: the reference CPU implementation of the path on this box's host cores (rank 0 only)"""
    if rank != 0:
        return
    import torch
    from nfc_laboratory_b200 import synth
    cores = host_cores()
    S = max(cores, min(4 * cores, 64))
    n = min(args.samples, 10_000_000)
    iq = synth.synth_batch(args.workload, S, n, seed=args.seed, device="cpu").numpy()
    for _ in range(args.warmup):
        cpu_reference_run(iq[:cores, : n // 8], cores)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        v, fr = cpu_reference_run(iq, cores)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libnfcref.so missing"}))
            return
        frames += fr
    dt = time.perf_counter() - t0
    value = S * n * args.steps / dt / 1e6
    sample = "%d streams x %d samples of the %s workload per step, %d host threads, IQ->magnitude included" % (S, n, args.workload, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.streams, args.samples, args.streams * args.samples * 8), "sample": sample,
                   "streams_per_gpu": args.streams, "samples_per_stream": args.samples, "sample_rate": RATE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_step": frames // max(1, args.steps),
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.quick:
        args.streams, args.samples = 64, 2_000_000

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import synth, dist as ND

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")

    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    flow_test = os.environ.get("NFCB200_BENCH_FLOW_TEST") == "1"   # tests/test_bench_contract.py: this function with a stand-in decoder, tensors on the host
    if flow_test:
        dev = "cpu"
    if world > 1:
        if flow_test:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device(dev))

    S, n = args.streams, args.samples
    free, total_mem = torch.cuda.mem_get_info()
    need = S * n * 8
    if need > 0.8 * free:
        S = max(1, int(0.8 * free // (n * 8)))
    bytes_per_step = S * n * 8

    # ---- synthetic batch, generated on the device (seed differs per rank: every rank holds different streams) -----------
    iq = torch.empty((S, n, 2), dtype=torch.float32, device=dev)
    synth.synth_batch(args.workload, S, n, seed=args.seed + 1000 * rank, device=dev, out=iq)
    torch.cuda.synchronize()

    dec = N.NfcDecoder(device=local, exact=args.exact)
    cap = max(1 << 16, S * (n // 12000 + 64))

    def step_device():
        buf, nf = dec.decode_batch_ptr(iq.data_ptr(), True, N.SIG_IQ_F32, S, n, RATE, cap=cap, raw=True)
        return buf, nf

    gather_ms = {"gather_pack": [], "gather_nccl": [], "gather_d2h": [], "gather_d2h_alloc": [], "gather_d2h_issue": [], "gather_d2h_wait": []}

    def gather(buf, nf):
        """frames of all ranks on rank 0.  GPUs: the packed device records of the decode go to rank 0 over NCCL point to point
        (no host round trip on the senders); CPU flow test: the host wire format over gloo"""
        if world == 1:
            return nf
        if flow_test or not hasattr(dec, "device_frames"):
            flat = ND.pack_frames(ND.frames_as_array(buf, nf), stream_offset=rank * S)
            allf = ND.gather_frames(flat, dev)
            return nf if allf is None else ND.count_frames(allf)
        tm = {}
        g = ND.gather_device_frames(dec, dev, lambda r: r * S, RATE, timings=tm)
        for k in gather_ms:
            gather_ms[k].append(tm.get(k, 0.0))
        return nf if g is None else nf + g.count

    parity = None  # set by the cpu_baseline leg (the only place this arm touches the oracle)

    # ---- warm-up, then the timed region -----------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 0)):
        buf, nf = step_device()
        gather(buf, nf)

    sampler = ClockSampler(local)
    sampler.start()  # every rank watches its own GPU: the slowest rank sets the step time

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stats = []
    frames_total = 0
    t_decode, t_gather = [], []
    for _ in range(args.steps):
        ta = time.perf_counter()
        buf, nf = step_device()
        tb = time.perf_counter()
        stats.append(dec.stats())
        frames_total += gather(buf, nf)
        t_decode.append((tb - ta) * 1e3)
        t_gather.append((time.perf_counter() - tb) * 1e3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0

    # digest of the last resident decode (frames of this rank): the host-input decode below must reproduce it
    digest_resident = ND.frames_digest(ND.frames_as_array(buf, nf))
    frames_last = ND.frames_as_array(buf, nf).copy() if (rank == 0 and world == 1 and not args.no_cpu and not args.no_parity) else None

    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())

    clocks = sampler.summary()

    # host wall time of the decode call and of the gather per step: slowest and fastest rank (the weak-scaling loss is the spread)
    rank_spread = None
    if world > 1:
        v = torch.tensor([statistics.mean(t_decode), statistics.mean(t_gather)], dtype=torch.float64, device=dev)
        vmax, vmin = v.clone(), v.clone()
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        rank_spread = {"decode_call_ms_max": float(vmax[0]), "decode_call_ms_min": float(vmin[0]), "gather_call_ms_max": float(vmax[1]),
                       "gather_call_ms_min": float(vmin[1])}
        mine = {"rank": rank, "decode_call_ms": statistics.mean(t_decode), "gather_call_ms": statistics.mean(t_gather),
                "sm_mhz": (clocks or {}).get("sm_mhz"), "power_w_max": (clocks or {}).get("power_w_max"), "reasons": (clocks or {}).get("reasons")}
        for k in ("ms_screen", "ms_segment", "ms_lanes", "ms_gather", "ms_total", "ms_wall"):
            mine[k] = statistics.mean(s_[k] for s_ in stats)
        mine["lanes"] = stats[-1]["lanes"]
        mine["lane_samples"] = stats[-1]["lane_samples"]
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        rank_spread["per_rank"] = per_rank

    # frame gather phases of the timed steps, max over ranks (the last gather_steps entries: warm-up and e2e gathers excluded)
    gather_phases = {}
    if world > 1 and gather_ms["gather_nccl"]:
        keys = list(gather_ms)
        vals = torch.tensor([statistics.mean(gather_ms[k][len(gather_ms[k]) - args.steps:]) for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        gather_phases = {k: float(v) for k, v in zip(keys, vals)}

    value = world * S * n * args.steps / dt / 1e6

    # ---- rooflines (HBM-bound path, algorithmic bytes = one read of the IQ input) ------------------------------
    # `roofline` is the DOMINANT kernel's: the lane kernel (K2), which only has to touch the active share of the samples
    # (lane_samples / samples of the input bytes); K1 (the dense screen that reads every byte) and the whole step follow.
    ms_screen = statistics.mean(s["ms_screen"] for s in stats)
    ms_lanes = statistics.mean(s["ms_lanes"] for s in stats)
    ms_step = statistics.mean(s["ms_total"] for s in stats)
    lane_frac = statistics.mean(s["lane_samples"] / max(1, s["samples"]) for s in stats)
    peak, peak_src = measured_peaks()
    achieved = bytes_per_step / (ms_screen * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "screen_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                t = json.load(f)
            # bytes per sample measured by ncu (dram read + write) scaled to this launch
            traffic = float(t["dram_bytes_per_sample"]) * S * n
        except Exception:
            traffic = None
    lanes_bytes = bytes_per_step * lane_frac
    lanes_achieved = lanes_bytes / (ms_lanes * 1e-3) / 1e9 if ms_lanes > 0 else 0.0
    step_achieved = bytes_per_step / (ms_step * 1e-3) / 1e9 if ms_step > 0 else 0.0
    lane_kernel = "wlanes_kernel (K2, warp lanes: one warp per stream, rings in shared memory)" if args.exact else \
        "lanes_kernel (K2, thread lanes: one thread per segment group, exact per-sample decoder)"
    roofline = {"bound": "hbm", "achieved": lanes_achieved, "peak": peak, "unit": "GB/s", "frac": lanes_achieved / peak, "traffic": lanes_traffic(args.exact, bytes_per_step),
                "kernel": lane_kernel, "peak_source": peak_src, "algorithmic_bytes_per_launch": lanes_bytes, "ms_per_launch": ms_lanes,
                "note": "algorithmic bytes = the active share of the input the lanes must read (%.3f of %d bytes); all lane + chain launches of a step" % (lane_frac, bytes_per_step)}
    roofline_k1 = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                   "kernel": "screen_kernel (K1: IQ->magnitude, prefix-sum moving sums, A/F/V correlators, B edge IIR)", "peak_source": peak_src,
                   "algorithmic_bytes_per_launch": bytes_per_step, "ms_per_launch": ms_screen}
    roofline_step = {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                     "algorithmic_bytes_per_launch": bytes_per_step, "ms_per_launch": ms_step, "kernel": "whole step (K1 + segments + lanes + chain + gather)"}

    # ---- e2e: the same decode through the C ABI with HOST buffers (H2D inside the timed region) ---------------------------
    e2e = None
    if not args.no_e2e:
        # every rank pins its own copy of (a part of) its batch: the whole batch on one GPU when it fits 42 % of the host
        # memory the job may use (cgroup limit), 12 GB per rank when several ranks share the host (PCIe-bound either way)
        avail = host_memory_budget()
        limit = min(0.42 * avail, 96 << 30) if world == 1 else min(0.25 * avail / world, 12 << 30)
        Se = S
        while Se > 1 and Se * n * 8 > limit:
            Se //= 2
        host = torch.empty((Se, n, 2), dtype=torch.float32, pin_memory=not flow_test)
        host.copy_(iq[:Se])
        torch.cuda.synchronize()
        dec.decode_batch_ptr(host.data_ptr(), False, N.SIG_IQ_F32, Se, n, RATE, cap=cap, raw=True)  # warm
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        d2h = 0
        esteps = 2 if (args.quick or flow_test) else max(5, min(args.steps, 8))  # at least five timed steps
        for _ in range(esteps):
            buf, nf = dec.decode_batch_ptr(host.data_ptr(), False, N.SIG_IQ_F32, Se, n, RATE, cap=cap, raw=True)
            d2h = nf * 128 + 8
            gather(buf, nf)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        de = time.perf_counter() - t1
        tm = torch.tensor([de], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        de = float(tm.item())
        same = None
        if Se == S:
            same = ND.frames_digest(ND.frames_as_array(buf, nf)) == digest_resident
            if not same:
                raise SystemExit("the host-input decode and the device-resident decode of the same batch differ: refusing to report a number")
        # the same batch as 16-bit IQ (the WAV ingest format, SIG_IQ_S16: 4 bytes per sample): half the PCIe bytes
        e2e_s16 = None
        try:
            # the int16 batch lives in the first half of the SAME page-locked buffer (no second 40 GB allocation: the caching
            # host allocator would keep the float buffer as well)
            host16 = host.view(torch.int16).reshape(-1)[: Se * n * 2].view(Se, n, 2)
            for c0 in range(0, Se, 16):
                host16[c0:c0 + 16].copy_((iq[c0:c0 + 16] * 32768.0).round().clamp_(-32768, 32767).to(torch.int16))
            torch.cuda.synchronize()
            if not flow_test:
                dec.decode_batch_ptr(host16.data_ptr(), False, N.SIG_IQ_S16, Se, n, RATE, cap=cap, raw=True)  # warm
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                nf16 = 0
                steps16 = min(esteps, 3)
                for _ in range(steps16):
                    buf16, nf16 = dec.decode_batch_ptr(host16.data_ptr(), False, N.SIG_IQ_S16, Se, n, RATE, cap=cap, raw=True)
                    gather(buf16, nf16)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                d16 = time.perf_counter() - t2
                tm16 = torch.tensor([d16], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tm16, op=dist.ReduceOp.MAX)
                d16 = float(tm16.item())
                e2e_s16 = {"value": world * Se * n * steps16 / d16 / 1e6, "unit": UNIT, "h2d_bytes_per_step": Se * n * 4, "streams": Se, "steps": steps16,
                           "frames_per_step": int(nf16), "note": "host-pinned int16 IQ (SIG_IQ_S16) -> nfcb200_decode_batch -> frames in host memory"}
            del host16
        except Exception as e:
            e2e_s16 = {"error": "%s: %s" % (type(e).__name__, e)}
        e2e = {"value": world * Se * n * esteps / de / 1e6, "unit": UNIT, "h2d_bytes_per_step": Se * n * 8, "d2h_bytes_per_step": int(d2h), "steps": esteps, "s16": e2e_s16,
               "streams": Se, "same_frames_as_resident": same, "note": "host-pinned float2 IQ -> nfcb200_decode_batch -> frames in host memory"
                                      + ("" if Se == S else " (sub-batch of %d streams: host memory bound)" % Se)}
        del host

    # ---- CPU baseline: the reference decoder on this box's cores, bounded sample of the same batch -------------------------
    cpu = None
    fullcheck = None
    fullparity = None
    wavset = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        Sc = min(S, max(cores, min(2 * cores, 32)))
        nc = min(n, 10_000_000)
        sub = iq[:Sc, :nc].cpu().numpy()
        v, fr = cpu_reference_run(sub, cores)
        if v is not None and not args.no_parity:
            # the same leg also checks the GPU arm against the reference (untimed): first streams, first 2e6 samples,
            # every frame field; a mismatch voids the run
            import nfcutil as U
            ns = min(2, Sc)
            m = min(nc, 2_000_000)
            got = dec.decode_batch(iq[:ns, :m].contiguous(), N.SIG_IQ_F32, RATE)
            parity = True
            for s_ in range(ns):
                mag = np.empty(m, dtype=np.float32)
                U.ref_lib().nfcref_iq_magnitude(np.ascontiguousarray(sub[s_, :m]).ctypes.data, m, mag.ctypes.data)
                if [f.key() for f in got if f.stream == s_] != U.ref_decode(mag, RATE):
                    parity = False
            if not parity:
                raise SystemExit("parity check against the reference oracle FAILED: refusing to report a number")
        if v is not None and frames_last is not None:
            fullcheck = full_size_check(frames_last, S, n, args.workload, args.seed + 1000 * rank, iq)
        if v is not None and frames_last is not None and not args.no_full_parity:
            fullparity = full_parity(frames_last, iq, S, n, cores, budget_s=args.parity_budget)
        if v is not None and not args.no_wav_set:
            wavset = wav_set(lambda exact: N.NfcDecoder(device=local, exact=exact), cores)
        if v is not None:
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%d of the batch's streams x %d samples, one NfcDecoder per host thread on %d threads, IQ->magnitude included" % (Sc, nc, cores)}

    if rank == 0:
        st = stats[-1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, S, n, bytes_per_step),
                       "streams_per_gpu": S, "samples_per_stream": n, "sample_rate": RATE, "sharding": "streams, block partition, NCCL frame gather (packed device records, point to point to rank 0)" if world > 1 else "single GPU",
                       "lanes": "warp lanes, one per stream, exact float state" if args.exact else "thread lanes, cold-started running sums"},
            "e2e": e2e,
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
            "roofline": roofline, "roofline_k1": roofline_k1, "roofline_step": roofline_step,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "phases_ms": dict({k: statistics.mean(s[k] for s in stats) for k in ("ms_screen", "ms_segment", "ms_lanes", "ms_gather", "ms_total", "ms_wall")},
                              **gather_phases),
            "decode": {"frames_per_step": frames_total // max(1, args.steps), "segments": st["segments"], "lanes": st["lanes"], "rounds": st["rounds"],
                       "lane_runs": st["lane_runs"], "lane_samples_frac": st["lane_samples"] / max(1, st["samples"])},
            "parity_spot_check": parity, "frames_digest": "%016x" % digest_resident, "full_size_check": fullcheck,
            "full_parity": fullparity,
            "wav_set": wavset,
            "rank_spread": rank_spread,
        }
        print(json.dumps(line))

    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
