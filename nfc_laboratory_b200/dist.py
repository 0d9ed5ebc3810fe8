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


def gather_frames(flat, device, group=None):
    """variable-length gather of every rank's packed frames to rank 0 (padded all_gather; the volume is O(frames), tiny
    next to the sample data, so this is latency- not bandwidth-bound).  Returns the concatenated buffer on rank 0, None
    elsewhere.  Works with NCCL (device='cuda:i') and gloo (device='cpu')."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    size = torch.tensor([flat.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if flat.size:
        mine[:flat.size] = torch.from_numpy(flat).to(device)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    if rank != 0:
        return None
    return np.concatenate([b[:s].cpu().numpy() for b, s in zip(bufs, sizes)])
