#!/usr/bin/env python3
"""BASELINE.json configs[4] shape on the GPUs of one box: ONE continuous synthetic NFC-A 424 kbps capture, time-sharded
across the ranks (one process per GPU) with the inter-rank carry exchange (dist.decode_long_capture_carry), frames gathered
to rank 0 and compared with the uncut decode of the same capture on one GPU.

launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/long_capture_bench.py --samples 200000000
(the full 8e9-sample capture of the config does not fit one process's uncut reference decode; the sample clock of the frame
format is 32 bits, so parity is defined below 2^32 samples anyway: SURVEY.md 8d)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200_000_000)
    ap.add_argument("--workload", default="nfca424")
    ap.add_argument("--seed", type=int, default=77)
    ap.add_argument("--left", type=int, default=8192)
    ap.add_argument("--overlap", type=int, default=1 << 18)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import synth, dist as ND

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    n = args.samples
    # every rank renders the same capture (same seed) and uses its own window of it
    iq = torch.empty((1, n, 2), dtype=torch.float32, device=dev)
    synth.synth_batch(args.workload, 1, n, seed=args.seed, device=dev, out=iq, chunk_streams=1)
    torch.cuda.synchronize()

    dec = N.NfcDecoder(device=local)
    window = lambda b, e: iq[:, b:e].contiguous()
    # warm-up (allocations) on a small window
    dec.decode_batch(window(0, min(n, 4_000_000)), N.SIG_IQ_F32, 10_000_000, cap=1 << 18)

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = {}
    frames = ND.decode_long_capture_carry(dec, window, n, world, N.SIG_IQ_F32, 10_000_000, overlap=args.overlap, left=args.left,
                                          rank=(rank if world > 1 else None), device=dev, stats=st)
    torch.cuda.synchronize()
    counts = torch.tensor([len(frames), st.get("redecoded", 0), st.get("redecoded_samples", 0)], dtype=torch.int64, device=dev)
    if world > 1:
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        dist.barrier()
    else:
        allc = [counts]
    dt = time.perf_counter() - t0
    tm = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dt = float(tm.item())

    # parity gate: rank 0 decodes the whole capture uncut; every rank's frames (hashed) travel to rank 0
    import numpy as np
    import zlib
    mine = [zlib.crc32(repr(f).encode()) for f in frames]
    h = torch.tensor(mine if mine else [0], dtype=torch.int64, device=dev)
    if world > 1:
        sizes = [int(c[0].item()) for c in allc]
        cap = max(max(sizes), 1)
        pad = torch.zeros(cap, dtype=torch.int64, device=dev)
        pad[:len(mine)] = h[:len(mine)]
        bufs = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(bufs, pad)
        stitched = [int(v) for b_, s_ in zip(bufs, sizes) for v in b_[:s_].tolist()]
    else:
        stitched = mine
    if rank == 0:
        t1 = time.perf_counter()
        full = dec.decode_batch(iq, N.SIG_IQ_F32, 10_000_000, cap=1 << 20)
        t_full = time.perf_counter() - t1
        ref = [zlib.crc32(repr((f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start, f.sample_end, f.data)).encode())
               for f in full]
        same = stitched == ref
        print(json.dumps({
            "metric": "IQ MSamples/s decoded", "config": {"workload": "%s: one continuous synthetic 10 MS/s float2 IQ capture of %d samples, time-sharded over %d GPU(s), "
                                                           "left overlap %d, right overlap %d, carry exchange rank to rank" % (args.workload, n, world, args.left, args.overlap)},
            "value": n / dt / 1e6, "unit": "MSamples/s", "n_gpus": world, "seconds": dt, "frames": len(stitched), "uncut_frames": len(ref),
            "stitched_equals_uncut": bool(same), "shards_decoded_from_an_injected_carry": int(sum(int(c[1].item()) for c in allc)),
            "samples_decoded_again": int(sum(int(c[2].item()) for c in allc)),
            "uncut_one_gpu_msps": n / t_full / 1e6}))
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
