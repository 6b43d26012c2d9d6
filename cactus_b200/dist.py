"""Multi-GPU plumbing: ends are independent, so the path shards with no data-path collective. What crosses ranks is
only (1) the end list dealt out by rank 0 (scatter) and (2) timing / cell counts / alignment checksums coming back
(all-reduce, gather). Backend-agnostic (nccl on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def deal_end_ranges(total_ends, world):
    """contiguous, balanced [first, count) ranges; rank r gets ranges[r]"""
    base, extra = divmod(total_ends, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def scatter_end_ranges(ranges, device):
    """rank 0 passes the list of (first, count); every rank receives its own pair"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return ranges[0]
    mine = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank() == 0:
        parts = [torch.tensor(list(r), dtype=torch.int64, device=device) for r in ranges]
        dist.scatter(mine, parts, src=0)
    else:
        dist.scatter(mine, None, src=0)
    return int(mine[0].item()), int(mine[1].item())


def reduce_stats(values, device):
    """values: list of floats -> (max over ranks, sum over ranks) as python lists"""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t.tolist(), t.tolist()
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.tolist(), sm.tolist()


def gather_checksums(value, device):
    """gather one float per rank on rank 0 (None elsewhere)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(value)]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())] if dist.get_rank() == 0 else None
    dist.gather(t, out, dst=0)
    return [float(x[0]) for x in out] if out is not None else None


_PINNED = {}


def _pinned(n, key):
    """grow-only pinned host buffer (plain memory when CUDA is absent, i.e. in the gloo tests)"""
    t = _PINNED.get(key)
    if t is None or t.numel() < n:
        t = torch.empty(max(n, 1) + max(n, 1) // 4, dtype=torch.uint8)
        if torch.cuda.is_available():
            t = t.pin_memory()
        _PINNED[key] = t
    return t[:n]


def gather_bytes(host_bytes, device):
    """gather a variable-length uint8 tensor per rank on rank 0: lengths first (all_gather of one int64), then every rank SENDS its
    un-padded payload and rank 0 RECEIVES them into consecutive slices of ONE device buffer (batched isend / irecv: an
    ncclSend / ncclRecv group on the GPU box), followed by one copy into pinned host memory. Returns the list of host tensors
    (views of that buffer) on rank 0, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [host_bytes]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([host_bytes.numel()], dtype=torch.int64, device=device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    lens = [int(x.item()) for x in lens]
    mine = host_bytes.to(device, non_blocking=True)
    if rank != 0:
        if lens[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, 0)]):     # batched like the receiver's side (one ncclGroup)
                w.wait()
        return None
    offs = [0]
    for v in lens:
        offs.append(offs[-1] + v)
    buf = torch.empty(max(offs[-1], 1), dtype=torch.uint8, device=device)
    buf[: lens[0]].copy_(mine)
    ops = [dist.P2POp(dist.irecv, buf[offs[r]: offs[r + 1]], r) for r in range(1, world) if lens[r]]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    host = _pinned(offs[-1], "gather")
    host.copy_(buf[: offs[-1]])
    return [host[offs[r]: offs[r + 1]] for r in range(world)]


def gather_msa_bytes(outs, msa_len, K, device):
    """the MSA rows of this rank's ends (malloc'd K x msa_len blocks, `outs[i]`) -> one pinned host buffer -> rank 0 (gather_bytes)"""
    import ctypes as C
    import numpy as np
    sizes = np.ascontiguousarray(msa_len.astype(np.int64) * K)
    total = int(sizes.sum())
    buf = _pinned(total, "msa")
    from .api import load_library
    load_library().barb200_pack_rows(outs, sizes.ctypes.data, len(sizes), buf.data_ptr())      # parallel copy into the pinned block
    return gather_bytes(buf, device)
