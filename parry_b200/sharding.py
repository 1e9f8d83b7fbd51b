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


class OverlappedHitGather:
    """Range-split ray batch whose per-rank result all-gather overlaps the traversal: the local shard is cast in
    `chunks` pieces on the compute stream, and as soon as piece c is done its (toi, tri) slices are all-gathered on a
    side stream while piece c + 1 is being traversed. Only the last piece's gather is exposed. Buffers are allocated
    once. `full()` returns rank-major (world * m_local) views of the gathered results.

    cast_fn(lo, hi, toi_out, tri_out) must enqueue the cast of local rays [lo, hi) on `compute_stream` (CUDA) or run it
    synchronously (CPU / gloo tests)."""

    def __init__(self, m_local, device, chunks=4, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.m = int(m_local)
        chunks = max(1, min(int(chunks), self.m)) if self.m else 1
        base, rem = divmod(self.m, chunks)
        self.spans, lo = [], 0
        for c in range(chunks):
            hi = lo + base + (1 if c < rem else 0)
            self.spans.append((lo, hi))
            lo = hi
        self.toi = torch.empty(self.m, dtype=torch.float32, device=device)
        self.tri = torch.empty(self.m, dtype=torch.int32, device=device)
        self.g_toi = [torch.empty(self.world * (hi - lo), dtype=torch.float32, device=device) for lo, hi in self.spans]
        self.g_tri = [torch.empty(self.world * (hi - lo), dtype=torch.int32, device=device) for lo, hi in self.spans]
        self.cuda = torch.device(device).type == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None
        self.events = [torch.cuda.Event() for _ in self.spans] if self.cuda else None

    def run(self, cast_fn, compute_stream=None):
        for c, (lo, hi) in enumerate(self.spans):
            cast_fn(lo, hi, self.toi[lo:hi], self.tri[lo:hi])
            if self.cuda:
                self.events[c].record(compute_stream)
                self.comm.wait_event(self.events[c])
                with torch.cuda.stream(self.comm):
                    dist.all_gather_into_tensor(self.g_toi[c], self.toi[lo:hi], group=self.group)
                    dist.all_gather_into_tensor(self.g_tri[c], self.tri[lo:hi], group=self.group)
            else:
                dist.all_gather_into_tensor(self.g_toi[c], self.toi[lo:hi].contiguous(), group=self.group)
                dist.all_gather_into_tensor(self.g_tri[c], self.tri[lo:hi].contiguous(), group=self.group)
        if self.cuda:
            (compute_stream or torch.cuda.current_stream()).wait_stream(self.comm)

    def full(self):
        """(toi, tri) of all ranks, rank-major: element r * m_local + i is ray i of rank r (all ranks hold m_local rays)."""
        toi = torch.cat([g.view(self.world, -1) for g in self.g_toi], dim=1).reshape(-1)
        tri = torch.cat([g.view(self.world, -1) for g in self.g_tri], dim=1).reshape(-1)
        return toi, tri


class PeerHitGather:
    """Same job as OverlappedHitGather, without NCCL kernels: the gather buffers are symmetric memory (every rank's buffer
    is mapped into every process), and the library pushes each finished piece of the local shard into all peers' buffers
    with copy-engine transfers over NVLink while the next piece is traversed (pb2_trimesh_cast_rays_allgather). A persistent
    traversal kernel fills every SM, so an NCCL all-gather kernel only gets to run once the traversal is over; DMA copies
    overlap for real. `run` ends with a cross-rank barrier on the compute stream. Raises if symmetric memory is unavailable
    (callers fall back to OverlappedHitGather)."""

    def __init__(self, m_local, device, chunks=4, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.m = int(m_local)
        self.chunks = int(chunks)
        n = self.world * self.m
        self.toi = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.tri = symm_mem.empty(n, dtype=torch.int32, device=device)
        self.h_toi = symm_mem.rendezvous(self.toi, self.group)
        self.h_tri = symm_mem.rendezvous(self.tri, self.group)
        self.p_toi = (C.c_void_p * self.world)(*[int(p) for p in self.h_toi.buffer_ptrs])
        self.p_tri = (C.c_void_p * self.world)(*[int(p) for p in self.h_tri.buffer_ptrs])

    def run(self, mesh, rays, max_toi, compute_stream):
        mesh.cast_local_ray_allgather(rays, max_toi, self.p_toi, self.p_tri, self.rank, self.rank * self.m, self.chunks)
        with torch.cuda.stream(compute_stream):
            self.h_toi.barrier()

    def local(self):
        lo = self.rank * self.m
        return self.toi[lo:lo + self.m], self.tri[lo:lo + self.m]

    def full(self):
        return self.toi, self.tri


class _DevView:
    """__cuda_array_interface__ over a raw device pointer (memory owned by the library), so that torch can view it."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": typestr, "version": 2}


class CommHitGather:
    """PeerHitGather on the C ABI alone: the gather buffers come from pb2_comm_peer_alloc (CUDA IPC mappings of every rank's
    buffer), the closing barrier is pb2_comm_barrier. Nothing here needs torch.distributed besides the one-off hand-over of the
    128-byte NCCL id (parry_b200.Comm.from_torch_distributed); a Rust host does the same through include/parry_b200.h."""

    def __init__(self, comm, m_local, device, chunks=4):
        self.comm, self.world, self.rank = comm, comm.nranks, comm.rank
        self.m, self.chunks = int(m_local), int(chunks)
        n = self.world * self.m
        self.p_toi = comm.peer_alloc(n * 4)
        self.p_tri = comm.peer_alloc(n * 4)
        self.toi = torch.as_tensor(_DevView(self.p_toi[self.rank], n, "<f4"), device=device)
        self.tri = torch.as_tensor(_DevView(self.p_tri[self.rank], n, "<i4"), device=device)

    def run(self, mesh, rays, max_toi, compute_stream=None):
        mesh.cast_local_ray_allgather(rays, max_toi, self.p_toi, self.p_tri, self.rank, self.rank * self.m, self.chunks)
        self.comm.barrier()

    def local(self):
        lo = self.rank * self.m
        return self.toi[lo:lo + self.m], self.tri[lo:lo + self.m]

    def full(self):
        return self.toi, self.tri
