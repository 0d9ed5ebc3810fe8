"""One long capture, time-sharded with a block-overlap stitch (BASELINE.json configs[4]).

CPU part: the stitch is exercised with the compiled reference decoder and with the host build of the lane pipeline as the
per-shard decoder (every shard cold-starts at its window, like one GPU per shard would), single process and over gloo.
GPU part: the same through the C ABI.  The bar is the full, uncut decode of the same capture, field for field."""
import os
import sys

import numpy as np
import pytest

import nfcutil as U
import screen_ref as S
from test_golden_oracle import committed_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = U.fixture_names()
OVERLAP = 400_000  # longer than the longest exchange of the regression captures (an NFC-V listen frame of 337 k samples)

needs_ref = pytest.mark.skipif(U.ref_lib() is None, reason="oracle/_ref/libnfcref.so not built")


def test_time_shards_cover_the_capture():
    from nfc_laboratory_b200 import dist as ND
    for n in (1, 255, 256, 100_000, 8_000_000_000):
        for g in (1, 2, 3, 8):
            sh = ND.time_shards(n, g, 1 << 20)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(sh[i][1] == sh[i + 1][0] for i in range(g - 1))
            assert all(wb <= b and e <= we and wb % 256 == 0 and b % 256 == 0 for b, e, wb, we in sh)


@needs_ref
@pytest.mark.parametrize("name", NAMES)
def test_stitch_with_the_reference_decoder(name):
    """every shard decoded by the reference itself, cold-started at the shard window: stitched == uncut"""
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav(name)
    full = committed_ref(name)[0]
    for shards in (2, 3):
        got = ND.decode_long_capture(lambda w: U.ref_decode(np.ascontiguousarray(w), rate), mag, shards, OVERLAP)
        assert got == full, (name, shards)


@needs_ref
def test_short_overlap_loses_only_frames_longer_than_the_overlap():
    """the documented limit of the stitch: a frame longer than the overlap is lost at a cut, nothing else changes"""
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav("test_NFC-V_26kbps_001")
    full = committed_ref("test_NFC-V_26kbps_001")[0]
    got = ND.decode_long_capture(lambda w: U.ref_decode(np.ascontiguousarray(w), rate), mag, 2, 131072)
    missing = [f for f in full if f not in got]
    assert [f for f in got if f not in full] == []
    assert len(missing) == 1 and missing[0][6] - missing[0][5] > 131072


@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_424kbps_001", "test_NFC-A_424kbps_002", "test_NFC-B_106kbps_002", "test_NFC-F_212kbps_003",
                                  "test_NFC-V_26kbps_001", "test_POLL_ABF_001"])
def test_stitch_with_the_lane_pipeline(name):
    """the host build of the device pipeline (screen model + speculative lanes) as the per-shard decoder"""
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav(name)
    full = committed_ref(name)[0]

    def decode(w):
        w = np.ascontiguousarray(w)
        return U.sim_pipeline(w, S.block_flags_device_model(w, S.ScreenParams(rate)), rate)[0]

    assert ND.decode_long_capture(decode, mag, 3, OVERLAP) == full


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import nfcutil as V
    from nfc_laboratory_b200 import dist as ND
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mag, rate, _ = V.fixture_wav(name)
    mine = ND.decode_long_capture(lambda w: V.ref_decode(np.ascontiguousarray(w), rate), mag, world, OVERLAP, rank=rank)
    a = np.zeros(len(mine), dtype=ND.FRAME_DTYPE)
    for i, f in enumerate(mine):
        a[i]["tech_type"], a[i]["frame_type"], a[i]["frame_flags"], a[i]["frame_phase"], a[i]["frame_rate"] = f[0], f[1], f[2], f[3], f[4]
        a[i]["sample_start"], a[i]["sample_end"], a[i]["length"] = f[5], f[6], len(f[7])
        a[i]["data"][: len(f[7])] = np.frombuffer(f[7], dtype=np.uint8)
    allf = ND.gather_frames(ND.pack_frames(a), "cpu")
    if rank == 0:
        q.put([f[1:] for f in ND.unpack_frames(allf)])
    dist.barrier()
    dist.destroy_process_group()


@needs_ref
def test_time_sharded_capture_over_gloo_world2():
    """one process per shard, frame gather to rank 0: the concatenation in rank order is the uncut decode"""
    import torch.multiprocessing as mp
    name = "test_NFC-A_106kbps_424kbps_001"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == committed_ref(name)[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_NFC-A_106kbps_424kbps_001", "test_NFC-F_212kbps_003", "test_NFC-V_26kbps_001"])
def test_time_sharded_capture_on_the_device(decoder, name):
    """shards decoded one after the other by the CUDA path (each call cold-starts at its window)"""
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav(name)
    full = committed_ref(name)[0]

    def decode(w):
        return [f.key() for f in decoder.decode_batch(np.ascontiguousarray(w)[None, :], N.SIG_MAG_F32, rate)]

    for shards in (2, 4):
        assert ND.decode_long_capture(decode, mag, shards, OVERLAP) == full, (name, shards)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_carry_exchange_makes_a_short_overlap_exact(name):
    """time shards with only 8 192 samples of left overlap: the plain overlap stitch loses frames on most captures (a cut
    between a poll frame and its answer, protocol state set before the overlap); with the decoder's carry handed from shard
    to shard at an idle point (dist.decode_long_capture_carry, NfcDecoder.carry_before / set_carry) the stitched decode
    equals the uncut one"""
    import nfc_laboratory_b200 as N
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav(name)
    d = N.NfcDecoder()
    full = [f.key() for f in d.decode_batch(mag[None], N.SIG_MAG_F32, rate)]
    for ranks, shards in ((False, 2), (False, 4), (True, 3), (True, 7)):
        st = {}
        got = ND.decode_long_capture_carry(d, lambda b, e: mag[None, b:e], mag.size, shards, N.SIG_MAG_F32, rate, overlap=OVERLAP, left=8192, stats=st,
                                           model_ranks=ranks, step=1 << 16 if shards == 7 else 1 << 20)
        assert got == full, (name, shards, ranks, st)
    d.close()


@pytest.mark.parametrize("ranks", [False, True], ids=["serial", "rank-protocol"])
@pytest.mark.parametrize("name", NAMES)
def test_carry_exchange_with_the_lane_pipeline(name, ranks):
    """dist.decode_long_capture_carry over the host build of the lane pipeline (nfcutil.HostWindowDecoder: the same segment /
    lane / carry-chain code as the device): with 8 192 samples of left overlap the stitched decode equals the uncut one.
    serial: every shard decoded once from the injected carry; rank-protocol: cold decodes first, then the re-run from the
    received carry until it meets the cold decode again (what one process per GPU does)"""
    from nfc_laboratory_b200 import dist as ND
    mag, rate, _ = U.fixture_wav(name)
    full = committed_ref(name)[0]
    d = U.HostWindowDecoder(rate)
    for shards in ((3, 7) if ranks else (2, 4)):
        st = {}
        got = ND.decode_long_capture_carry(d, lambda b, e: mag[None, b:e], mag.size, shards, None, rate, overlap=OVERLAP, left=8192, stats=st, model_ranks=ranks,
                                           step=1 << 16 if shards == 7 else 1 << 20)
        assert got == full, (name, shards, st)


def _carry_worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import nfcutil as V
    from nfc_laboratory_b200 import dist as ND
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mag, rate, _ = V.fixture_wav(name)
    st = {}
    mine = ND.decode_long_capture_carry(V.HostWindowDecoder(rate), lambda b, e: mag[None, b:e], mag.size, world, None, rate, overlap=OVERLAP, left=8192,
                                        rank=rank, device="cpu", stats=st)
    out = [None] * world
    dist.all_gather_object(out, (mine, st))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("test_NFC-A_106kbps_001", 3), ("test_NFC-F_212kbps_001", 2), ("test_NFC-B_106kbps_002", 2)])
def test_carry_exchange_over_gloo(name, world):
    """one process per shard: cold decodes in parallel, the carries sent rank to rank (send / recv of a byte tensor), the
    bounds all-gathered; the concatenation in rank order is the uncut decode"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_carry_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    frames = [f for part, _ in got for f in part]
    assert frames == committed_ref(name)[0], [st for _, st in got]


def test_rank_protocol_rejoins_the_cold_decode():
    """a synthetic 424 kbps capture of 1.2 s on two shards: the second shard's cold decode misses the session state, the
    re-run from the received carry meets it again once the saturating NFC-F pulse counter of the cold decode has caught up
    (canonical carry: every value from 94 on is one state) -- the rest of the shard is NOT decoded again"""
    from nfc_laboratory_b200 import synth, dist as ND
    n = 12_000_000
    iq = synth.synth_batch("nfca424", 1, n, seed=77, device="cpu")[0].numpy()
    mag = np.sqrt(iq[:, 0].astype(np.float32) ** 2 + iq[:, 1].astype(np.float32) ** 2).astype(np.float32)
    d = U.HostWindowDecoder(10_000_000)
    key = lambda f: (f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start, f.sample_end, f.data)
    full = [key(f) for f in d.decode_batch(mag[None])]
    st = {}
    got = ND.decode_long_capture_carry(d, lambda b, e: mag[None, b:e], n, 2, None, 10_000_000, overlap=1 << 18, left=8192, stats=st, model_ranks=True,
                                       step=1 << 18)
    assert got == full
    assert st["redecoded"] == 1 and st["redecoded_samples"] < 0.8 * (n // 2), st
