"""Multi-GPU plumbing for the query batches: every ray / query box / candidate pair is an independent
unit, so batches are range-split across ranks with the Bvh and shape tables replicated per GPU and NO data-path
collective. The only exchanges are the result gathers below (torch.distributed: NCCL on GPUs, gloo in the CPU tests):
fixed-size per-ray hit records use one all_gather; variable-size compacted lists (pairs, contacts) gather their counts
first and then a padded all_gather."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of a batch of n units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def interleaved_ids(n, rank, world):
    """Interleaved ownership i = rank (mod world), used for per-leaf self-pair walks (balances the j > i filter)."""
    return torch.arange(rank, n, world)


def all_gather_hits(toi, tri, n_total, group=None):
    """Gathers per-ray (toi, tri) shards into full-size arrays on every rank. Shards may differ by one element."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    rec = torch.zeros((pad, 2), dtype=torch.int32, device=toi.device)
    lo, hi = sizes[rank]
    rec[: hi - lo, 0] = toi.view(torch.int32)
    rec[: hi - lo, 1] = tri.view(torch.int32) if tri.dtype != torch.int32 else tri
    out = torch.empty((world * pad, 2), dtype=torch.int32, device=toi.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    out = out.view(world, pad, 2)
    full = torch.cat([out[r, : sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    return full[:, 0].contiguous().view(torch.float32), full[:, 1].contiguous()


def all_gather_counts(count, device, group=None):
    """All-gather of one per-rank count (the 'compacted hit/pair counts' collective)."""
    world = dist.get_world_size(group)
    mine = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out


def all_gather_varlen(rows, group=None):
    """Variable-size gather of compacted records (rows: (count, k) tensor): counts first, then a padded all_gather."""
    world = dist.get_world_size(group)
    counts = all_gather_counts(rows.shape[0], rows.device, group)
    pad = int(counts.max().item())
    k = rows.shape[1]
    buf = torch.zeros((pad, k), dtype=rows.dtype, device=rows.device)
    buf[: rows.shape[0]] = rows
    out = torch.empty((world * pad, k), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, pad, k)
    return torch.cat([out[r, : int(counts[r])] for r in range(world)], dim=0), counts
