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


def pack_frames(arr, stream_offset=0, chunk=65536):
    """frame records -> one flat uint8 buffer: per frame an 80-byte header followed by `length` payload bytes"""
    parts = []
    for i in range(0, arr.size, chunk):
        a = arr[i:i + chunk]
        head = np.zeros((a.size, HEADER_BYTES), dtype=np.uint8)
        hv = head.view(np.dtype([("u", "<u4", (8,)), ("q", "<u8", (3,)), ("d", "<f8", (3,))]))[:, 0]
        hv["u"][:, 0] = a["stream"] + stream_offset
        for k, name in enumerate(("tech_type", "frame_type", "frame_flags", "frame_phase", "frame_rate", "length")):
            hv["u"][:, k + 1] = a[name]
        hv["q"][:, 0] = a["sample_start"]
        hv["q"][:, 1] = a["sample_end"]
        hv["q"][:, 2] = a["sample_rate"]
        hv["d"][:, 0] = a["time_start"]
        hv["d"][:, 1] = a["time_end"]
        hv["d"][:, 2] = a["date_time"]
        ln = a["length"].astype(np.int64)
        mask = np.arange(512)[None, :] < ln[:, None]
        width = HEADER_BYTES + ln
        offs = np.concatenate([[0], np.cumsum(width)])
        flat = np.empty(int(offs[-1]), dtype=np.uint8)
        hidx = offs[:-1, None] + np.arange(HEADER_BYTES)[None, :]
        flat[hidx.ravel()] = head.ravel()
        pidx = (offs[:-1, None] + HEADER_BYTES + np.arange(512)[None, :])[mask]
        flat[pidx] = a["data"][mask]
        parts.append(flat)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


def unpack_frames(flat):
    """inverse of pack_frames -> list of (stream, tech, type, flags, phase, rate, start, end, payload bytes)"""
    out = []
    pos = 0
    n = flat.size
    while pos < n:
        u = flat[pos:pos + 32].view("<u4")
        q = flat[pos + 32:pos + 56].view("<u8")
        ln = int(u[6])
        out.append((int(u[0]), int(u[1]), int(u[2]), int(u[3]), int(u[4]), int(u[5]), int(q[0]), int(q[1]),
                    bytes(flat[pos + HEADER_BYTES:pos + HEADER_BYTES + ln])))
        pos += HEADER_BYTES + ln
    return out


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
