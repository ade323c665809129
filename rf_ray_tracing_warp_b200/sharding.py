"""Multi-GPU plumbing: rays shard by contiguous global ray-id range, the mesh/BVH is replicated, and the only
exchange step of the data path is ONE all-gather of fixed-size record segments (include/rfrt.h: the counts and the
job's counters ride in the segment headers, so the exchange needs no host round trip).  Works on any
torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests)."""
import numpy as np
import torch

SEG_HEADER_U64 = 16
CTR_COUNT = 10  # include/rfrt.h RFRT_CTR_COUNT


def ray_range(n_rays, rank, world):
    """Contiguous share of global ray ids [0, n_rays) for `rank` of `world` (covers every id exactly once)."""
    return (rank * n_rays // world, (rank + 1) * n_rays // world)


def exchange_segments(all_segments, local_segment, group=None):
    """all_segments (world * nbytes uint8) <- every rank's local_segment (nbytes uint8), in rank order."""
    dist = torch.distributed
    try:
        dist.all_gather_into_tensor(all_segments, local_segment, group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        world = dist.get_world_size(group)
        parts = list(all_segments.view(world, -1).unbind(0))
        dist.all_gather(parts, local_segment, group=group)
    return all_segments


def _align16(x):
    return (x + 15) & ~15


def segment_layout(capacity, path_floats):
    """Byte offsets of the sections of a record segment — the host mirror of the layout in include/rfrt.h
    (checked against rfrt_record_segment_bytes by the tests)."""
    off = {}
    pos = 8 * SEG_HEADER_U64
    for name, width in (("ray", 4), ("rx", 4), ("nverts", 4), ("bin", 8), ("amp", 8), ("dist", 8)):
        off[name] = pos
        pos += _align16(width * capacity)
    off["paths"] = pos
    pos += _align16(4 * capacity * path_floats)
    off["total"] = pos
    return off


_SECTION_DTYPES = dict(ray=np.uint32, rx=np.int32, nverts=np.int32, bin=np.int64, amp=np.float64, dist=np.float64)


def read_segments(buffer, n_segments, capacity, path_floats):
    """Host-side reader of `n_segments` consecutive segments (uint8 array / CPU tensor): list of dicts with the header
    words and the stored records.  Debug / test helper — the product sorts segments on the device."""
    raw = np.asarray(buffer.cpu() if torch.is_tensor(buffer) else buffer, dtype=np.uint8).reshape(n_segments, -1)
    lay = segment_layout(capacity, path_floats)
    out = []
    for s in range(n_segments):
        seg = raw[s]
        header = seg[: 8 * SEG_HEADER_U64].view(np.uint64)
        n = int(min(header[0], header[1]))
        d = dict(produced=int(header[0]), fit=int(header[1]), counters=header[2:2 + CTR_COUNT].copy())
        for name, dt in _SECTION_DTYPES.items():
            d[name] = seg[lay[name]: lay[name] + n * np.dtype(dt).itemsize].view(dt).copy()
        if path_floats:
            d["paths"] = seg[lay["paths"]: lay["paths"] + 4 * n * path_floats].view(np.float32).reshape(n, path_floats).copy()
        out.append(d)
    return out


def write_segment(records, counters, capacity, path_floats):
    """Host-side writer (test helper; the product packs on the device with rfrt_records_pack)."""
    lay = segment_layout(capacity, path_floats)
    seg = np.zeros(lay["total"], dtype=np.uint8)
    produced = len(records["ray"])
    n = min(produced, capacity)
    header = seg[: 8 * SEG_HEADER_U64].view(np.uint64)
    header[0], header[1] = produced, capacity
    header[2:2 + CTR_COUNT] = np.asarray(counters, dtype=np.uint64)
    for name, dt in _SECTION_DTYPES.items():
        seg[lay[name]: lay[name] + n * np.dtype(dt).itemsize] = np.ascontiguousarray(records[name][:n], dtype=dt).view(np.uint8)
    if path_floats:
        seg[lay["paths"]: lay["paths"] + 4 * n * path_floats] = \
            np.ascontiguousarray(records["paths"][:n], dtype=np.float32).reshape(-1).view(np.uint8)
    return seg


def sum_stats(stats, device, group=None):
    keys = sorted(stats)
    t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=device)
    torch.distributed.all_reduce(t, group=group)
    return dict(zip(keys, t.cpu().tolist()))
