"""world_size-2 gloo run of the multi-GPU plumbing (cactus_b200/dist.py) on CPU: end ranges dealt by rank 0,
max/sum reductions and the checksum gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cactus_b200 import dist as D
    dev = torch.device("cpu")
    ranges = D.deal_end_ranges(101, world) if rank == 0 else None
    first, n = D.scatter_end_ranges(ranges, dev)
    mx, sm = D.reduce_stats([10.0 + rank, float(n)], dev)
    ck = D.gather_checksums(first * 1000 + n, dev)
    gb = D.gather_bytes(torch.arange(3 + 2 * rank, dtype=torch.uint8) + 10 * rank, dev)
    q.put((rank, first, n, mx, sm, ck, None if gb is None else [t.tolist() for t in gb]))
    dist.destroy_process_group()


def test_scatter_reduce_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, mx0, sm0, ck0, gb0), (r1, f1, n1, mx1, sm1, ck1, gb1) = res
    assert gb0 == [[0, 1, 2], [10, 11, 12, 13, 14]] and gb1 is None
    assert (f0, n0) == (0, 51) and (f1, n1) == (51, 50)
    assert mx0 == mx1 == [11.0, 51.0] and sm0 == sm1 == [21.0, 101.0]
    assert ck0 == [51.0, 51050.0] and ck1 is None


def test_deal_is_a_partition():
    from cactus_b200 import dist as D
    for total, world in [(0, 1), (7, 8), (100000, 8), (2368, 3)]:
        r = D.deal_end_ranges(total, world)
        assert r[0][0] == 0 and sum(n for _, n in r) == total
        assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))
