"""Multi-GPU plumbing: rays shard by contiguous global ray-id range, the mesh/BVH is replicated, and the only
exchange step is gathering the (sparse) received records.  Works on any torch.distributed backend (NCCL on
the GPUs; gloo in the CPU tests)."""
import torch


def ray_range(n_rays, rank, world):
    """Contiguous share of global ray ids [0, n_rays) for `rank` of `world` (covers every id exactly once)."""
    return (rank * n_rays // world, (rank + 1) * n_rays // world)


def gather_records(rec, group=None):
    """All-gather variable-length record arrays (dict name -> tensor with the same leading dim, or None).
    Every rank returns the concatenation over ranks in rank order."""
    dist = torch.distributed
    world = dist.get_world_size(group)
    first = next(v for v in rec.values() if v is not None)
    n_local = torch.tensor([first.shape[0]], dtype=torch.int64, device=first.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(x.item()) for x in counts]
    m = max(max(counts), 1)
    out = {}
    for k, t in rec.items():
        if t is None:
            out[k] = None
            continue
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out


def sort_records(rec):
    """(receiver, ray id) order == the reference's accumulation order per receiver (tracer.py:87,102)."""
    key = (rec["rx"].to(torch.int64) << 32) | (rec["ray"].to(torch.int64) & 0xFFFFFFFF)
    order = torch.argsort(key)
    return {k: (v[order].contiguous() if v is not None else None) for k, v in rec.items()}


def sum_stats(stats, device, group=None):
    keys = sorted(stats)
    t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=device)
    torch.distributed.all_reduce(t, group=group)
    return dict(zip(keys, t.cpu().tolist()))
