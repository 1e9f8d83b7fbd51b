#!/usr/bin/env python
"""Multi-rank check of the C ABI's exchange layer (pb2_comm_*, pb2_bvh_self_pairs_shard, peer buffers for
pb2_trimesh_cast_rays_allgather). Launch with one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 harness/comm_check.py
torch.distributed (gloo) is used only to carry the 128-byte NCCL id from rank 0 to the others and to cross-check results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FMAX = float(np.finfo(np.float32).max)


def main():
    import torch
    import torch.distributed as dist
    import parry_b200
    from harness import scenes
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = parry_b200.Context(local)
    stream = ctx.torch_stream()
    comm = parry_b200.Comm.from_torch_distributed(ctx)
    assert comm.rank == rank and comm.nranks == world

    # 1. counts
    counts = comm.allgather_counts(100 + 7 * rank)
    assert counts == [100 + 7 * r for r in range(world)], counts

    # 2. variable-size gather of compacted rows, against the rows every rank can regenerate
    def rows_of(r):
        g = scenes.rng(1000 + r)
        k = int(g.integers(0, 5000)) if r != 1 else 0          # one rank contributes nothing
        return g.integers(0, 1 << 30, (k, 3)).astype(np.int32)
    with torch.cuda.stream(stream):
        mine = torch.from_numpy(rows_of(rank)).cuda()
    ctx.synchronize()
    allrows, cnts = comm.allgatherv(mine, capacity=16)           # too small on purpose: the retry path is collective
    ctx.synchronize()
    want = np.concatenate([rows_of(r) for r in range(world)])
    assert cnts == [len(rows_of(r)) for r in range(world)]
    assert (allrows.cpu().numpy() == want).all()

    # 3. broad phase split over the ranks: shards partition the pair set
    n = 60000
    kinds, params, poses, _ = scenes.colliders(n, seed=5)
    shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) for k, p in zip(kinds, params)])
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    full = np.asarray(bvh.traverse_bvtt_single_tree()).astype(np.int64)
    with torch.cuda.stream(stream):
        d_aabbs = torch.from_numpy(aabbs).cuda()
    ctx.synchronize()
    part = bvh.traverse_bvtt_single_tree_shard(rank, world, like=d_aabbs)
    gathered, pc = comm.allgatherv(part)
    ctx.synchronize()
    key = lambda p: np.sort(np.asarray(p).astype(np.int64) @ np.array([1 << 32, 1]))
    gk = key(gathered.cpu().numpy().view(np.uint32))
    assert len(gk) == len(full) and (gk == key(full)).all(), (len(gk), len(full))
    assert len(np.unique(gk)) == len(gk)
    assert max(pc) < 1.3 * (len(full) / world) + 64, pc             # interleaved ownership balances the shards

    # 4. ray shards gathered through peer buffers (copy-engine pushes while the traversal kernel runs)
    v, i = scenes.uv_sphere(96, 64)
    mesh = parry_b200.TriMesh(ctx, v, i)
    m = 1 << 16
    rays = [scenes.sphere_rays(m, seed=50 + r) for r in range(world)]
    p_toi, p_tri = comm.peer_alloc(world * m * 4), comm.peer_alloc(world * m * 4)
    with torch.cuda.stream(stream):
        d_rays = torch.from_numpy(rays[rank]).cuda()
    ctx.synchronize()
    mesh.cast_local_ray_allgather(d_rays, FMAX, p_toi, p_tri, rank, rank * m, 4)
    comm.barrier()
    ctx.synchronize()
    import ctypes as C
    got_toi, got_tri = np.empty(world * m, np.float32), np.empty(world * m, np.uint32)
    cudart = C.CDLL("libcudart.so.12")        # already loaded by torch
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert cudart.cudaMemcpy(got_toi.ctypes.data, p_toi[rank], world * m * 4, 2) == 0
    assert cudart.cudaMemcpy(got_tri.ctypes.data, p_tri[rank], world * m * 4, 2) == 0
    for r in range(world):
        t, k = mesh.cast_local_ray(rays[r], FMAX)
        assert (got_toi[r * m:(r + 1) * m].view(np.uint32) == np.asarray(t).view(np.uint32)).all(), "toi shard %d" % r
        assert (got_tri[r * m:(r + 1) * m] == np.asarray(k)).all(), "tri shard %d" % r

    # 5. fixed-size all-gather in place
    with torch.cuda.stream(stream):
        buf = torch.zeros(world * 1024, dtype=torch.int32, device="cuda")
        buf[rank * 1024:(rank + 1) * 1024] = rank + 1
    ctx.synchronize()
    comm.allgather(buf[rank * 1024:(rank + 1) * 1024], buf)
    ctx.synchronize()
    assert (buf.view(world, 1024).cpu().numpy() == (np.arange(world) + 1)[:, None]).all()
    comm.close()
    dist.barrier()
    if rank == 0:
        print("comm_check ok: %d ranks, %d pairs in %s-pair shards, %d rays per rank gathered through peer buffers" % (world, len(full), pc, m))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
