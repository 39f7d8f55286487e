"""world_size-2 gloo tests of the multi-GPU plumbing (request sharding, one-shot weight broadcast, result gather)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT / "riffusion-hobby_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from riffusion import distributed as rd

    spec = [("a.weight", (3, 5)), ("a.bias", (3,)), ("b.weight", (2, 2, 3, 3))]
    sd = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        sd = {n: torch.randn(s, generator=g).half() for n, s in spec}
    got = rd.broadcast_state_dict(spec, sd, src=0, device="cpu")
    checksum = float(sum(t.float().sum() for t in got.values()))
    n_items = 7
    mine = rd.shard_range(n_items, rank, world)
    local = torch.tensor([[i, i * i] for i in mine], dtype=torch.float32)       # "results" of this rank's requests
    allres = rd.gather_results(local, n_items)
    q.put((rank, checksum, list(mine), None if allres is None else allres.tolist(),
           {n: tuple(t.shape) for n, t in got.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_shard_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, c0, m0, all0, shapes0), (r1, c1, m1, all1, shapes1) = res
    assert c0 == c1 and shapes0 == shapes1 == {"a.weight": (3, 5), "a.bias": (3,), "b.weight": (2, 2, 3, 3)}
    assert m0 == [0, 1, 2, 3] and m1 == [4, 5, 6]                   # contiguous blocks, first rank takes the remainder
    assert all1 is None and all0 == [[i, i * i] for i in range(7)]


def test_shard_range_covers_everything():
    sys.path.insert(0, str(ROOT / "riffusion-hobby_b200"))
    from riffusion.distributed import shard_range

    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            items = [i for r in range(world) for i in shard_range(n, r, world)]
            assert items == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
