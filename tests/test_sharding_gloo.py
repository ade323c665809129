"""CPU test of the multi-rank host logic with the gloo backend, world_size 2: ray-range sharding covers every
ray id once, the segment exchange (one all-gather of fixed-size record segments) leaves the same bytes on every rank,
and the segments merged in (receiver, ray id) order equal the single-rank records."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(begin, end, B=3):
    """Deterministic stand-in for the device records of ray ids [begin, end): 1 ray in 7 is 'received'."""
    ids = np.arange(begin, end, dtype=np.int64)
    ids = ids[(ids * 2654435761 % 7) == 0]
    rng = np.random.default_rng(1234)
    table = rng.random(100003)
    return dict(ray=ids.astype(np.uint32), rx=(ids * 40503 % 5).astype(np.int32), nverts=(2 + ids % 3).astype(np.int32),
                bin=(ids % 977).astype(np.int64), amp=table[ids % 100003], dist=table[(ids * 3) % 100003] * 50.0,
                paths=np.repeat(ids[:, None].astype(np.float32), (B + 1) * 3, axis=1))


def _merge(segments):
    """NumPy statement of what rfrt_records_sort does with the gathered segments: (receiver, ray id) order."""
    cat = {k: np.concatenate([s[k] for s in segments]) for k in ("ray", "rx", "nverts", "bin", "amp", "dist", "paths")}
    order = np.argsort((cat["rx"].astype(np.int64) << 32) | cat["ray"].astype(np.int64), kind="stable")
    return {k: v[order] for k, v in cat.items()}


CAP, ROW = 9000, 12


def _worker(rank, world, port, n_rays, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from rf_ray_tracing_warp_b200 import sharding
    begin, end = sharding.ray_range(n_rays, rank, world)
    counters = [end - begin, 11 * (rank + 1), 0, 5, 0, 0, 0, 0, 0, 0]
    local = torch.from_numpy(sharding.write_segment(_fake_records(begin, end), counters, CAP, ROW))
    gathered = torch.empty(world * local.numel(), dtype=torch.uint8)
    sharding.exchange_segments(gathered, local)
    stats = sharding.sum_stats(dict(segments=end - begin, rank_sum=rank), "cpu")
    torch.save(dict(gathered=gathered, stats=stats, range=(begin, end)), os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_segment_exchange(tmp_path):
    """Every rank ends up with the same gathered segments; merged in (receiver, ray id) order they equal the records of
    a single rank that traced all rays; the counters in the headers add up."""
    from rf_ray_tracing_warp_b200 import sharding
    n_rays, world = 100001, 2
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert outs[0]["range"][0] == 0 and outs[0]["range"][1] == outs[1]["range"][0] and outs[1]["range"][1] == n_rays
    assert torch.equal(outs[0]["gathered"], outs[1]["gathered"])
    single = _merge(sharding.read_segments(sharding.write_segment(_fake_records(0, n_rays), [0] * sharding.CTR_COUNT, 2 * CAP, ROW), 1, 2 * CAP, ROW))
    for o in outs:
        assert o["stats"]["segments"] == n_rays and o["stats"]["rank_sum"] == 1
        segs = sharding.read_segments(o["gathered"], world, CAP, ROW)
        assert [int(s["counters"][1]) for s in segs] == [11, 22]
        assert sum(int(s["counters"][0]) for s in segs) == n_rays
        assert all(s["produced"] <= s["fit"] for s in segs)
        merged = _merge(segs)
        for k, v in single.items():
            assert np.array_equal(merged[k], v), k
    key = single["rx"].astype(np.int64) * (1 << 32) + single["ray"].astype(np.int64)
    assert bool((key[1:] > key[:-1]).all())


def test_segment_overflow_is_visible():
    from rf_ray_tracing_warp_b200 import sharding
    rec = _fake_records(0, 5000)
    seg = sharding.read_segments(sharding.write_segment(rec, [0] * sharding.CTR_COUNT, 100, 0), 1, 100, 0)[0]
    assert seg["produced"] == len(rec["ray"]) > seg["fit"] == 100 and len(seg["ray"]) == 100


def test_segment_layout_matches_library():
    """The host mirror of the segment layout agrees with the C ABI's size query (host-only call, no GPU needed)."""
    import ctypes
    from rf_ray_tracing_warp_b200 import _lib, sharding
    lib = _lib.load()
    for cap, row in [(1, 0), (100, 0), (4097, 27), (1 << 18, 0), (1 << 16, 99)]:
        n = ctypes.c_int64(0)
        assert lib.rfrt_record_segment_bytes(cap, row, n) == 0
        assert n.value == sharding.segment_layout(cap, row)["total"]
    n = ctypes.c_int64(0)
    assert lib.rfrt_records_workspace_bytes(1 << 20, n) == 0 and n.value > 36 << 20
    assert lib.rfrt_record_segment_bytes(-1, 0, n) == -1


def test_ray_ranges_partition():
    from rf_ray_tracing_warp_b200 import sharding
    for n in [0, 1, 7, 1 << 20, (1 << 28) + 3]:
        for world in [1, 2, 3, 8]:
            r = [sharding.ray_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
