"""CPU test of the multi-rank host logic with the gloo backend, world_size 2: ray-range sharding covers every
ray id once, the record exchange returns the same (receiver, ray id)-ordered records on every rank, and that
result equals the single-rank one."""
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
    rx = (ids * 40503 % 5).astype(np.int32)
    return dict(ray=torch.from_numpy(ids.astype(np.int32)), rx=torch.from_numpy(rx),
                bin=torch.from_numpy((ids % 977).astype(np.int64)),
                amp=torch.from_numpy(table[ids % 100003]),
                paths=torch.from_numpy(np.repeat(ids[:, None, None].astype(np.float32), (B + 1) * 3, axis=1).reshape(-1, B + 1, 3)),
                none=None)


def _worker(rank, world, port, n_rays, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from rf_ray_tracing_warp_b200 import sharding
    begin, end = sharding.ray_range(n_rays, rank, world)
    rec = sharding.sort_records(sharding.gather_records(_fake_records(begin, end)))
    stats = sharding.sum_stats(dict(segments=end - begin, records=int(rec["ray"].shape[0]) if rank == 0 else 0), "cpu")
    torch.save(dict(rec=rec, stats=stats, range=(begin, end)), os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_record_exchange(tmp_path):
    from rf_ray_tracing_warp_b200 import sharding
    n_rays, world = 100001, 2
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert outs[0]["range"][0] == 0 and outs[0]["range"][1] == outs[1]["range"][0] and outs[1]["range"][1] == n_rays
    single = sharding.sort_records(_fake_records(0, n_rays))
    for o in outs:
        assert o["stats"]["segments"] == n_rays
        for k, v in single.items():
            if v is None:
                assert o["rec"][k] is None
            else:
                assert torch.equal(o["rec"][k], v), k
    key = single["rx"].to(torch.int64) * (1 << 32) + single["ray"].to(torch.int64)
    assert bool((key[1:] > key[:-1]).all())


def test_ray_ranges_partition():
    from rf_ray_tracing_warp_b200 import sharding
    for n in [0, 1, 7, 1 << 20, (1 << 28) + 3]:
        for world in [1, 2, 3, 8]:
            r = [sharding.ray_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
