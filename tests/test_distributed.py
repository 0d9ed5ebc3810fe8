"""N > 1 path on CPU: stream sharding and the variable-length frame gather over torch.distributed (gloo, world 2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_streams_partitions():
    from nfc_laboratory_b200 import dist as ND
    for n in (1, 7, 8, 1024, 8192):
        for world in (1, 2, 3, 8):
            parts = [ND.shard_streams(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _fake_frames(rng, n, stream_lo, stream_hi):
    from nfc_laboratory_b200 import dist as ND
    a = np.zeros(n, dtype=ND.FRAME_DTYPE)
    a["stream"] = np.sort(rng.integers(stream_lo, stream_hi, n))
    a["tech_type"] = 0x101
    a["frame_type"] = rng.choice([0x102, 0x103], n)
    a["frame_rate"] = 105938
    a["length"] = rng.integers(0, 300, n)
    a["sample_start"] = rng.integers(0, 1 << 31, n)
    a["sample_end"] = a["sample_start"] + 1000
    a["sample_rate"] = 10_000_000
    for i in range(n):
        a["data"][i, : a["length"][i]] = rng.integers(0, 256, a["length"][i])
    return a


def test_pack_unpack_roundtrip():
    from nfc_laboratory_b200 import dist as ND
    rng = np.random.default_rng(3)
    a = _fake_frames(rng, 500, 0, 50)
    flat = ND.pack_frames(a, stream_offset=7)
    out = ND.unpack_frames(flat)
    assert len(out) == 500
    for rec, f in zip(a, out):
        assert f[0] == rec["stream"] + 7 and f[2] == rec["frame_type"] and f[6] == rec["sample_start"]
        assert f[8] == bytes(rec["data"][: rec["length"]])
    assert ND.count_frames(flat) == 500
    empty = ND.pack_frames(a[:0])
    assert ND.unpack_frames(empty) == [] and ND.count_frames(np.concatenate([empty, flat, empty])) == 500


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from nfc_laboratory_b200 import dist as ND
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    lo, hi = ND.shard_streams(64, rank, world)
    a = _fake_frames(rng, 40 + 17 * rank, 0, hi - lo)
    flat = ND.pack_frames(a, stream_offset=lo)
    allf = ND.gather_frames(flat, "cpu")
    if rank == 0:
        frames = ND.unpack_frames(allf)
        q.put((len(frames), [f[0] for f in frames]))
    else:
        assert allf is None
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_frame_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n, streams = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n == 40 + 57
    # rank 0 owns streams [0, 32), rank 1 [32, 64): the gathered list is ordered by rank, hence by stream block
    assert streams[:40] == sorted(streams[:40]) and max(streams[:40]) < 32
    assert min(streams[40:]) >= 32
