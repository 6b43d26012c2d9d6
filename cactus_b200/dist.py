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


def gather_bytes(host_bytes, device):
    """gather a variable-length uint8 tensor per rank on rank 0 (device buffers over the backend's transport; returns
    the list of host tensors on rank 0, None elsewhere). Lengths travel first, payloads are padded to the longest."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [host_bytes]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([host_bytes.numel()], dtype=torch.int64, device=device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    mx = max(int(x.item()) for x in lens)
    pad = torch.zeros(mx, dtype=torch.uint8, device=device)
    pad[: host_bytes.numel()].copy_(host_bytes, non_blocking=True)
    out = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0)
    if rank != 0:
        return None
    return [out[r][: int(lens[r].item())].cpu() for r in range(world)]
