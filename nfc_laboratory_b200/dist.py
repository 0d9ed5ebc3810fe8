"""Multi-GPU plumbing: capture streams are independent, so a batch shards across ranks with no data-path collective;
the only exchange is the final variable-length frame gather to rank 0.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) is plumbing here, not the product.
"""
import numpy as np

FRAME_DTYPE = np.dtype([
    ("stream", "<u4"), ("tech_type", "<u4"), ("frame_type", "<u4"), ("frame_flags", "<u4"), ("frame_phase", "<u4"),
    ("frame_rate", "<u4"), ("length", "<u4"), ("reserved", "<u4"), ("sample_start", "<u8"), ("sample_end", "<u8"),
    ("sample_rate", "<u8"), ("time_start", "<f8"), ("time_end", "<f8"), ("date_time", "<f8"), ("data", "u1", (512,)),
])  # include/nfcb200.h nfcb200_frame

HEADER_BYTES = 80


def shard_streams(n_streams, rank, world):
    """contiguous block of streams owned by `rank` (block partition, remainder to the low ranks)"""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frames_as_array(cbuf, n):
    """zero-copy numpy view of a ctypes nfcb200_frame array"""
    if n == 0:
        return np.zeros(0, dtype=FRAME_DTYPE)
    return np.frombuffer(cbuf, dtype=FRAME_DTYPE, count=n)


def frames_digest(arr):
    """order-dependent 64-bit digest of a frame array (every header field the reference compares plus the payload bytes):
    two decodes of the same batch -- device-resident or from host memory, one GPU or several -- must agree on it"""
    a = np.asarray(arr)
    n = int(a.size)
    if n == 0:
        return 0
    mask = np.uint64(0xFFFFFFFFFFFFFFFF)
    h = np.zeros(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k, name in enumerate(("stream", "tech_type", "frame_type", "frame_flags", "frame_phase", "frame_rate", "length", "sample_start", "sample_end")):
            h = (h * np.uint64(0x100000001B3) + a[name].astype(np.uint64) + np.uint64(k + 1)) & mask
        L = int(a["length"].max())
        if L:
            d = a["data"][:, :L].astype(np.uint64)
            live = np.arange(L, dtype=np.uint64)[None, :] < a["length"].astype(np.uint64)[:, None]
            w = (np.arange(L, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) & mask
            h = (h + ((d + np.uint64(1)) * w[None, :] * live).sum(axis=1, dtype=np.uint64)) & mask
        idx = np.arange(1, n + 1, dtype=np.uint64)
        total = int(((h ^ (idx * np.uint64(0xD6E8FEB86659FD93))) * np.uint64(0xFF51AFD7ED558CCD)).sum(dtype=np.uint64))
    return total & 0xFFFFFFFFFFFFFFFF


def frame_hashes(arr):
    """one 64-bit hash per frame over every field RawFrame::operator== compares plus the payload (NOT the stream index):
    the value oracle/ref_wrap.cpp nfcref_hash_batch computes for the reference's frames (full-size differential)"""
    a = np.asarray(arr)
    n = int(a.size)
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFFFFFFFFFF)
    h = np.zeros(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k, name in enumerate(("tech_type", "frame_type", "frame_flags", "frame_phase", "frame_rate", "length", "sample_start", "sample_end")):
            h = (h * np.uint64(0x100000001B3) + a[name].astype(np.uint64) + np.uint64(k + 1)) & mask
        L = int(a["length"].max())
        if L:
            d = a["data"][:, :L].astype(np.uint64)
            live = np.arange(L, dtype=np.uint64)[None, :] < a["length"].astype(np.uint64)[:, None]
            w = (np.arange(L, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) & mask
            h = (h + ((d + np.uint64(1)) * w[None, :] * live).sum(axis=1, dtype=np.uint64)) & mask
    return h


def pack_frames(arr, stream_offset=0):
    """frame records (numpy FRAME_DTYPE array, e.g. a view of the decoder's output buffer) -> one flat uint8 buffer
    [u64 count][count x 80-byte headers][payload bytes back to back], packed by nfcb200_pack_frames (host threads)"""
    import ctypes as C
    from .binding import load_library, _check
    lib = load_library()
    a = np.ascontiguousarray(arr)
    n = int(a.size)
    need = C.c_uint64(0)
    ptr = a.ctypes.data if n else None
    _check(lib, lib.nfcb200_pack_frames(ptr, n, int(stream_offset), None, 0, C.byref(need)))
    flat = np.empty(int(need.value), dtype=np.uint8)
    _check(lib, lib.nfcb200_pack_frames(ptr, n, int(stream_offset), flat.ctypes.data, flat.size, C.byref(need)))
    return flat


def unpack_frames(flat):
    """inverse of pack_frames (also accepts several packed buffers back to back, as gather_frames returns them)
    -> list of (stream, tech, type, flags, phase, rate, start, end, payload bytes)"""
    out = []
    pos = 0
    while pos < flat.size:
        n = int(flat[pos:pos + 8].view("<u8")[0])
        head = flat[pos + 8:pos + 8 + n * HEADER_BYTES].reshape(n, HEADER_BYTES)
        pay = pos + 8 + n * HEADER_BYTES
        for i in range(n):
            u = head[i, :32].view("<u4")
            q = head[i, 32:56].view("<u8")
            ln = int(u[6])
            out.append((int(u[0]), int(u[1]), int(u[2]), int(u[3]), int(u[4]), int(u[5]), int(q[0]), int(q[1]), bytes(flat[pay:pay + ln])))
            pay += ln
        pos = pay
    return out


def count_frames(flat):
    """number of frames in a (possibly concatenated) packed buffer without unpacking the payloads"""
    total = 0
    pos = 0
    while pos < flat.size:
        n = int(flat[pos:pos + 8].view("<u8")[0])
        head = flat[pos + 8:pos + 8 + n * HEADER_BYTES].reshape(n, HEADER_BYTES)
        total += n
        pos += 8 + n * HEADER_BYTES + int(head[:, 24:28].copy().view("<u4").sum())
    return total


_pinned = {}


def _pinned_bytes(key, n):
    """a cached page-locked uint8 host tensor of at least n bytes (None when pinning is not possible, e.g. no CUDA)"""
    import torch
    import os
    import sys
    import time
    try:
        t0 = time.perf_counter()
        t = _pinned.get(key)
        fresh = t is None or t.numel() < n
        if fresh:
            t = torch.empty(max(n, 1) + max(n, 1) // 4, dtype=torch.uint8, pin_memory=True)
            _pinned[key] = t
        if os.environ.get("NFCB200_GATHER_DEBUG"):
            print("[dist] pinned %s: need %d, have %d, %s, %.2f ms" % (key, n, t.numel(), "allocated" if fresh else "cached", (time.perf_counter() - t0) * 1e3),
                  file=sys.stderr, flush=True)
        return t
    except Exception as e:
        if os.environ.get("NFCB200_GATHER_DEBUG"):
            print("[dist] pinned %s failed: %s" % (key, e), file=sys.stderr, flush=True)
        return None


def gather_frames(flat, device, group=None):
    """variable-length gather of every rank's packed frames to rank 0 (padded all_gather; the volume is O(frames), tiny
    next to the sample data, so this is latency- not bandwidth-bound).  Returns the concatenated buffer on rank 0, None
    elsewhere.  Works with NCCL (device='cuda:i') and gloo (device='cpu').  On a GPU the host sides of the two copies go
    through cached page-locked buffers (a pageable copy of a few hundred MB would dominate the step at 8 ranks), and the
    returned array is a view of that buffer, valid until the next call."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    on_gpu = str(device).startswith("cuda")
    size = torch.tensor([flat.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if flat.size:
        src = torch.from_numpy(flat)
        stage = _pinned_bytes("send", flat.size) if on_gpu else None
        if stage is not None:
            stage[:flat.size].copy_(src)
            mine[:flat.size].copy_(stage[:flat.size], non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the staging buffer is reused by the next call
        else:
            mine[:flat.size] = src.to(device)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    if rank != 0:
        return None
    total = sum(sizes)
    stage = _pinned_bytes("recv", total) if on_gpu else None
    if stage is None:
        return np.concatenate([b[:s].cpu().numpy() for b, s in zip(bufs, sizes)])
    pos = 0
    for b, s in zip(bufs, sizes):
        if s:
            stage[pos:pos + s].copy_(b[:s], non_blocking=True)
        pos += s
    torch.cuda.synchronize()
    return stage[:total].numpy()  # a view of the cached staging buffer: valid until the next gather_frames call


class _DevView:
    """zero-copy torch view of raw device memory (torch.as_tensor reads __cuda_array_interface__)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_device_frames(dec, device, stream_offset, sample_rate, group=None, timings=None):
    """frame gather of the batch decode WITHOUT a host round trip on the sending side: every rank hands the device-resident,
    already ordered and packed records of its last decode (NfcDecoder.device_frames) to NCCL; one all_gather of the counts,
    then point-to-point transfers to rank 0 only.  Rank 0 copies what it received to page-locked host memory
    (GatheredRecords; .frames(dec) converts to ABI frames).  Returns None on the other ranks.  stream_offset(r) gives the
    first global stream index of rank r.  timings (dict) receives gather_pack / gather_nccl / gather_d2h in milliseconds."""
    import time
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t0 = time.perf_counter()
    rp, n, ep, ne = dec.device_frames()
    counts = torch.tensor([n, ne], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = [(int(c[0].item()), int(c[1].item())) for c in allc]
    t1 = time.perf_counter()

    if rank != 0:
        ops = []
        if n:
            ops.append(dist.P2POp(dist.isend, torch.as_tensor(_DevView(rp, n * 128), device=device), 0, group))
        if ne:
            ops.append(dist.P2POp(dist.isend, torch.as_tensor(_DevView(ep, ne * 128), device=device), 0, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        torch.cuda.current_stream().synchronize()
        if timings is not None:
            timings.update(gather_pack=(t1 - t0) * 1e3, gather_nccl=(time.perf_counter() - t1) * 1e3, gather_d2h=0.0)
        return None

    # page-locked staging for what the other ranks send (cached: allocated on the first call, grown when needed)
    t1a = time.perf_counter()
    tot_r = sum(c[0] for c in allc[1:])
    tot_e = sum(c[1] for c in allc[1:])
    t1x = time.perf_counter()
    hr = _pinned_bytes("recv_r", max(tot_r, 1) * 128)
    he = _pinned_bytes("recv_e", max(tot_e, 1) * 128)
    t1b = time.perf_counter()
    if __import__("os").environ.get("NFCB200_GATHER_DEBUG"):
        print("[dist] gather rank0: counts %.3f ms, sums %.3f ms, staging %.3f ms" % ((t1 - t0) * 1e3, (t1x - t1a) * 1e3, (t1b - t1x) * 1e3),
              file=__import__("sys").stderr, flush=True)

    bufs = []
    ops = []
    for r in range(1, world):
        nr, er = allc[r]
        rb = torch.empty(max(nr, 1) * 128, dtype=torch.uint8, device=device)
        eb = torch.empty(max(er, 1) * 128, dtype=torch.uint8, device=device)
        bufs.append((rb, eb))
        if nr:
            ops.append(dist.P2POp(dist.irecv, rb[:nr * 128], r, group))
        if er:
            ops.append(dist.P2POp(dist.irecv, eb[:er * 128], r, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()

    # device -> page-locked host, one staging buffer for all ranks' records and one for their extension chunks
    t2a = time.perf_counter()
    pr = pe = 0
    spans = []
    for (rb, eb), (nr, er) in zip(bufs, allc[1:]):
        if nr:
            hr[pr:pr + nr * 128].copy_(rb[:nr * 128], non_blocking=True)
        if er:
            he[pe:pe + er * 128].copy_(eb[:er * 128], non_blocking=True)
        spans.append((pr, nr, pe, er))
        pr += nr * 128
        pe += er * 128
    t2b = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t3 = time.perf_counter()

    if timings is not None:
        timings.update(gather_pack=(t1 - t0) * 1e3, gather_nccl=(t2 - t1b) * 1e3, gather_d2h=(t3 - t2) * 1e3,
                       gather_d2h_alloc=(t1b - t1a) * 1e3, gather_d2h_issue=(t2b - t2a) * 1e3, gather_d2h_wait=(t3 - t2b) * 1e3)
    return GatheredRecords(hr, he, spans, stream_offset, sample_rate)


class GatheredRecords:
    """what rank 0 holds after gather_device_frames: the other ranks' frames as packed 128-byte records (+ extension
    chunks) in page-locked host memory, one span per rank in rank order.  frames(dec) converts them to ABI frames."""

    def __init__(self, records, ext, spans, stream_offset, sample_rate):
        self.records, self.ext, self.spans = records, ext, spans
        self.stream_offset, self.sample_rate = stream_offset, sample_rate
        self.count = sum(s[1] for s in spans)

    def frames(self, dec):
        out = []
        for i, (pr, nr, pe, er) in enumerate(self.spans):
            if nr:
                out += dec.emit_records(self.records.data_ptr() + pr, nr, self.ext.data_ptr() + pe, er, self.stream_offset(i + 1), self.sample_rate)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# one long capture, time-sharded with a block-overlap stitch (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------------------------------
BLOCK = 256


def time_shards(n_samples, n_shards, overlap):
    """[(own_begin, own_end, window_begin, window_end)] of a capture cut into `n_shards` contiguous time shards.  A shard
    is decoded over its window -- its own samples plus `overlap` samples either side -- by a decoder cold-started at
    window_begin, and keeps the frames that START inside its own range.  `overlap` must exceed the longest exchange
    (poll frame + frame waiting time + listen frame): the frame a shard keeps then always has its poll frame inside the
    same window, which is the only context a cold-started decoder cannot re-derive (448 / 475 arbitrary
    cold starts reproduce the full decode field for field, every miss is a listen frame whose poll was cut off)."""
    cuts = [min(n_samples, (n_samples * r // n_shards) // BLOCK * BLOCK) for r in range(n_shards)] + [n_samples]
    out = []
    for r in range(n_shards):
        b, e = cuts[r], cuts[r + 1]
        out.append((b, e, max(0, b - overlap) // BLOCK * BLOCK, min(n_samples, e + overlap)))
    return out


def stitch_shard(frames, own_begin, own_end, window_begin, first):
    """frames of one shard window (tuples (tech, type, flags, phase, rate, start, end, payload), window-relative sample
    indices) -> the frames this shard owns, in absolute sample indices.  The first shard keeps everything before its end
    (including the reference's two carrier-off frames at samples 0 and 1); later shards drop what starts in their
    overlap, which also removes the cold-start artefacts of their decoder."""
    out = []
    for f in frames:
        start = f[5] + window_begin
        if (first or start >= own_begin) and start < own_end:
            out.append(f[:5] + (start, f[6] + window_begin) + tuple(f[7:]))
    return out


def decode_long_capture(decode, samples, n_shards, overlap=1 << 20, rank=None):
    """time-sharded decode of ONE capture: `decode(window) -> frames` is called once per shard (rank=None: all shards
    here, e.g. one GPU after the other; rank=r: only shard r, for one process per GPU followed by gather_frames).
    Exact as long as no sticky protocol state (FSD / FWT from RATS / ATTRIB, the Encrypted flag after AUTH) is set
    further back than `overlap` before a cut -- the block-overlap stitch BASELINE.json names, stated here as it is."""
    shards = time_shards(len(samples), n_shards, overlap)
    todo = range(n_shards) if rank is None else [rank]
    out = []
    for r in todo:
        b, e, wb, we = shards[r]
        if e <= b:
            continue
        out += stitch_shard(decode(samples[wb:we]), b, e, wb, r == 0)
    return out


NO_LANE = 0xFFFFFFFFFFFFFFFF
GAP = 32 * BLOCK  # an idle point (a lane's begin) has this many samples without activity in front of it (NFCB200_GAP_BLOCKS)


def _same_carry(a, base_a, b, base_b):
    """two carry blobs, each with the window base its clock counts from, describe the same decoder state"""
    if a[:-4] != b[:-4]:
        return False
    ea, eb = int.from_bytes(a[-4:], "little"), int.from_bytes(b[-4:], "little")  # carrier edge time, 0 = unset
    return (ea == 0 and eb == 0) or (ea != 0 and eb != 0 and ea + base_a == eb + base_b)


class _CarryShard:
    """One time shard of a capture on one decoder: cold decode, the carry received from the predecessor, the re-run of
    the decoder from the received carry up to the point where it agrees with the cold decode again, the answer for the
    successor.  A message is (blob, clock_base, L, covered): the carry in front of the idle point L (absolute sample, None
    when the sender has no idle point at or after the receiver's window start) with the base its clock counts from;
    covered: the sender's decode reaches the end of the capture, so without an idle point it owns the rest."""

    def __init__(self, dec, window, n_samples, shards, r, sigtype, sample_rate, lead, step, overlap):
        self.dec, self.window, self.n, self.shards, self.r, self.overlap = dec, window, n_samples, shards, r, overlap
        self.sigtype, self.rate, self.lead, self.step = sigtype, sample_rate, lead, step
        self.b, self.e, self.wb, self.we = shards[r]
        nxt = [q for q in range(r + 1, len(shards)) if shards[q][1] > shards[q][0]]
        self.next_wb = shards[nxt[0]][2] if nxt else None
        self.lo = 0
        self.kept = []         # [(frames with absolute times, from, to)]: what this shard keeps of each decode
        self.live_base = None  # window base of the decode the decoder handle holds
        self.live_end = 0
        self.next_answer = None
        self.relay = None
        self.decodes = 0       # decodes after the cold one
        self.samples = 0       # samples decoded after the cold one

    def _decode(self, b, e, carry=None, clock_base=0):
        self.dec.set_carry(carry, max(0, b - clock_base))
        fr = self.dec.decode_batch(self.window(b, e), self.sigtype, self.rate, cap=1 << 18)
        self.live_base, self.live_end = b, e
        return [(f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start + b, f.sample_end + b, f.data) for f in fr]

    def _query(self, sample):
        """(blob, clock base, idle point or None) of the decode in the handle: first lane that begins at or after `sample`"""
        blob, lane = self.dec.carry_before(max(0, sample - self.live_base))
        return blob, self.live_base, (None if lane is None else self.live_base + lane)

    def cold(self):
        self.cold_frames = self._decode(self.wb, self.we)
        self.kept = [(self.cold_frames, 0 if self.r == 0 else self.b, self.n)]

    def receive(self, msg):
        blob, cbase, L, covered = msg
        if L is None:
            if covered:
                self.lo, self.kept, self.relay = self.n, [], msg  # the predecessor's decode owns the rest of the capture
            else:
                # no idle point within the predecessor's reach: the plain overlap stitch, with its full overlap to the left
                self.lo = self.b
                nb = max(0, self.b - self.overlap) // BLOCK * BLOCK
                if nb < self.wb:
                    self.cold_frames = self._decode(nb, self.we)
                    self.decodes += 1
                    self.samples += self.we - nb
                self.kept = [(self.cold_frames, self.b, self.n)]
            return
        self.lo = L
        self.kept = [(self.cold_frames, L, self.n)]
        mine = self._query(L)
        # the successor's idle point must not lie before this shard's own: what precedes L is the predecessor's
        ask = max(self.next_wb, L) if self.next_wb is not None else None
        nq = self._query(ask) if ask is not None else None
        if mine[2] == L and _same_carry(mine[0], mine[1], blob, cbase):
            return  # the cold start assumed the right state
        # idle points of the cold decode further on: where the re-run may meet it again
        cands, s = [], self.step
        while L + s < self.we - GAP:
            c = self._query(L + s)
            if c[2] is None:
                break
            if c[2] > L and (not cands or c[2] > cands[-1][2]):
                cands.append(c)
            s = max(s * 4, c[2] - L + BLOCK)
        if nq is not None and nq[2] is not None and nq[2] > L and all(c[2] != nq[2] for c in cands):
            cands = sorted(cands + [nq], key=lambda c: c[2])
        pending_next = nq is not None
        start, carry = L, (blob, cbase)
        self.kept = []
        for c in cands + [None]:
            end = self.we if c is None else min(self.we, c[2] + GAP)
            nb = max(carry[1], (start - self.lead) // BLOCK * BLOCK, 0)
            frames = self._decode(nb, end, carry[0], carry[1])
            self.decodes += 1
            self.samples += end - nb
            if c is None:
                self.kept.append((frames, start, self.n))
                if pending_next:
                    self.next_answer = None  # answered from the handle: it holds the decode up to the window's end
                return
            here = self._query(c[2])
            if pending_next and ask >= start:
                a = self._query(ask)
                if a[2] is not None and a[2] <= c[2]:
                    self.next_answer, pending_next = a + (False,), False
            if here[2] != c[2]:
                continue  # the lanes of the re-run are cut differently here: try the next idle point with a longer window
            self.kept.append((frames, start, c[2]))
            if _same_carry(here[0], here[1], c[0], c[1]) and (not pending_next or nq[2] is None or nq[2] >= c[2]):
                self.kept.append((self.cold_frames, c[2], self.n))  # from here on the cold decode was right
                if pending_next:
                    self.next_answer = nq + (nq[2] is None and self.we >= self.n,)
                return
            start, carry = c[2], (here[0], here[1])

    def answer(self):
        """the message for the successor"""
        if self.relay is not None:
            return self.relay
        if self.next_answer is not None:
            return self.next_answer
        blob, base, L = self._query(max(self.next_wb, self.lo))
        return blob, base, L, (L is None and self.live_end >= self.n)

    def owned(self, hi):
        out = []
        for frames, lo, to in self.kept:
            lo, to = max(lo, self.lo), min(to, hi)
            out += [f for f in frames if lo <= f[5] < to]
        return out


def decode_long_capture_carry(dec, window, n_samples, n_shards, sigtype, sample_rate, overlap=1 << 18, left=8192, lead=6144, rank=None, group=None,
                              device=None, stats=None, model_ranks=False, step=1 << 20):
    """Time-sharded decode of ONE capture WITH the inter-shard carry exchange (SURVEY.md 8e "exchange step").

    The predecessor of a shard hands over the decoder's CARRY -- protocol state (FSD / FWT / SFGT from RATS / ATS / ATTRIB,
    the Encrypted flag, lastCommand), carrier flags, carrier edge time: NfcDecoder.carry_before -- in front of the first
    lane L that begins at or after the successor's window start (`left` samples before its own range): an idle point of
    the capture.  Frames that start before L belong to the predecessor, from L on to the successor.

    rank=None: all shards in this process in time order, every shard decoded ONCE from `lead` samples before its idle
    point with the carry injected (NfcDecoder.set_carry).
    rank=r: shard r of a torch.distributed group.  All ranks first decode their windows in parallel, cold.  Then the
    carries travel rank to rank (NCCL / gloo send-recv of a byte tensor).  A successor whose cold decode holds another state
    at L re-runs the decoder from L with the received carry -- in pieces that end at the idle points of its cold decode,
    and only until the state of the re-run equals the cold decode's again (then the rest of the cold decode was right).
    model_ranks=True runs that rank protocol for all shards in this process (one decoder, shard after shard).

    dec: nfc_laboratory_b200.NfcDecoder; window(b, e) -> samples [b, e) of the capture in `sigtype` layout (numpy array or
    CUDA tensor).  Returns this process's frames (absolute sample indices; all frames when rank is None).
    stats: "redecoded" (shards that ran from a received carry), "redecoded_samples".
    Not covered: a stretch of activity without any idle point (8 192 quiet samples) longer than the overlap -- there the
    plain overlap stitch of decode_long_capture is used (cold start `overlap` samples before the shard); a carrier edge older than the window stamps a later carrier frame with the window's start."""
    shards = [(b, e, max(0, b - left) // BLOCK * BLOCK, we) for (b, e, wb, we) in time_shards(n_samples, n_shards, overlap)]
    live = [r for r in range(n_shards) if shards[r][1] > shards[r][0]]

    def tup(f, base):
        return (f.tech_type, f.frame_type, f.frame_flags, f.frame_phase, f.frame_rate, f.sample_start + base, f.sample_end + base, f.data)

    if rank is None and not model_ranks:
        out, redecoded, resamples = [], 0, 0
        pending = None  # (frames, first owned sample) of the decode in the handle
        cur_base = cur_end = 0
        for i, r in enumerate(live):
            b, e, wb, we = shards[r]
            if i == 0:
                base, carry, lo = wb, None, 0
            else:
                blob, lane = dec.carry_before(max(0, max(wb, pending[1]) - cur_base))  # never before this decode's own idle point
                if lane is None:
                    if cur_end >= n_samples:
                        break                      # the decode in hand reaches the capture's end: it owns the rest
                    base, carry, lo = max(0, b - overlap) // BLOCK * BLOCK, None, b  # no idle point in reach: the plain overlap stitch
                else:
                    lo = cur_base + lane
                    base, carry = max(cur_base, (lo - lead) // BLOCK * BLOCK), blob
                    redecoded += 1
                    resamples += we - base
            if pending is not None:
                out += [f for f in pending[0] if pending[1] <= f[5] < lo]
            dec.set_carry(carry, base - cur_base)
            frames = [tup(f, base) for f in dec.decode_batch(window(base, we), sigtype, sample_rate, cap=1 << 18)]
            pending = (frames, lo)
            cur_base, cur_end = base, we
        if pending is not None:
            out += [f for f in pending[0] if pending[1] <= f[5]]
        if stats is not None:
            stats["redecoded"], stats["redecoded_samples"] = redecoded, resamples
        return out

    def worker(r):
        return _CarryShard(dec, window, n_samples, shards, r, sigtype, sample_rate, lead, step, overlap)

    if rank is None:
        # the rank protocol, shard after shard on one decoder
        ws, msg = [], None
        for i, r in enumerate(live):
            w = worker(r)
            w.cold()
            if i > 0:
                w.receive(msg)
            if i + 1 < len(live):
                msg = w.answer()
            ws.append(w)
        out = []
        for i, w in enumerate(ws):
            out += w.owned(ws[i + 1].lo if i + 1 < len(ws) else n_samples)
        if stats is not None:
            stats["redecoded"] = sum(1 for w in ws if w.decodes)
            stats["redecoded_samples"] = sum(w.samples for w in ws)
        return out

    import numpy as np
    import torch
    import torch.distributed as dist

    csize = dec.carry_size()
    mine = rank in live
    w = worker(rank) if mine else None
    if mine:
        w.cold()                                        # pass 1: every rank, in parallel
    lo = 0 if (not mine or rank == live[0]) else None
    for i in range(1, len(live)):
        src, dst = live[i - 1], live[i]
        if rank == src:
            blob, cbase, L, covered = w.answer()
            tail = np.array([NO_LANE if L is None else L, cbase, 1 if covered else 0], dtype="<u8").tobytes()
            t = torch.frombuffer(bytearray(blob + tail), dtype=torch.uint8).to(device)
            dist.send(t, dst, group=group)
        elif rank == dst:
            t = torch.empty(csize + 24, dtype=torch.uint8, device=device)
            dist.recv(t, src, group=group)
            payload = t.cpu().numpy().tobytes()
            tail = np.frombuffer(payload[csize:csize + 24], dtype="<u8")
            L = int(tail[0])
            w.receive((payload[:csize], int(tail[1]), None if L == NO_LANE else L, bool(tail[2])))
            lo = w.lo
    if stats is not None:
        stats["redecoded"] = 1 if (mine and w.decodes) else 0
        stats["redecoded_samples"] = w.samples if mine else 0
    # the bound towards the successor is known to the successor: fetch it
    bounds = torch.tensor([lo if mine else n_samples], dtype=torch.int64, device=device)
    allb = [torch.zeros_like(bounds) for _ in range(dist.get_world_size(group))]
    dist.all_gather(allb, bounds, group=group)
    if not mine:
        return []
    idx = live.index(rank)
    hi = int(allb[live[idx + 1]].item()) if idx + 1 < len(live) else n_samples
    return w.owned(hi)
