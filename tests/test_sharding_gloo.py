"""CPU, world_size 2 (gloo): the multi-GPU host logic of SURVEY §8e — range-split batches, replicated scene, result
all-gathers — reassembles exactly the single-process answer. The per-shard compute stand-in is the CPU oracle (the
GPU path itself is covered by the -m gpu tests; bench.py --gpus N uses the same sharding helpers over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMAX = float(np.finfo(np.float32).max)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from harness import oracle, scenes
        from parry_b200 import sharding
        # rays: contiguous range split, fixed-size records
        v, i = scenes.uv_sphere(32, 24)
        rays = scenes.sphere_rays(5001, seed=9)  # odd size: shards differ by one
        mesh = oracle.TriMesh(v, i)
        lo, hi = sharding.shard_range(len(rays), rank, world)
        toi, tri = mesh.cast_rays(None, rays[lo:hi], FMAX)
        ftoi, ftri = sharding.all_gather_hits(torch.from_numpy(toi), torch.from_numpy(tri.astype(np.int32)), len(rays))
        rtoi, rtri = mesh.cast_rays(None, rays, FMAX)
        ok_rays = bool((ftoi.numpy().view(np.uint32) == rtoi.view(np.uint32)).all() and (ftri.numpy().view(np.uint32) == rtri).all())
        # contacts: range split, variable-size compacted gather
        pts, radii = scenes.hull_pool(16)
        T = oracle.ShapeTable([("convex", p) for p in pts])
        a, b, p1, p2 = scenes.hull_pairs(3001, radii, seed=10)
        lo, hi = sharding.shard_range(len(a), rank, world)
        out, st = T.contact(a[lo:hi], p1[lo:hi], b[lo:hi], p2[lo:hi], 0.01)
        idx = np.nonzero(st == 1)[0]
        rows = np.concatenate([(idx + lo).astype(np.float64)[:, None], out[idx].astype(np.float64)], axis=1)
        allrows, counts = sharding.all_gather_varlen(torch.from_numpy(rows))
        rout, rst = T.contact(a, p1, b, p2, 0.01)
        ridx = np.nonzero(rst == 1)[0]
        got = allrows.numpy()
        ok_contacts = bool(int(counts.sum()) == len(ridx) and (got[:, 0].astype(np.int64) == ridx).all()
                           and (got[:, 1:].astype(np.float32).view(np.uint32) == rout[ridx].view(np.uint32)).all())
        # chunk-overlapped gather (same code path as bench.py --gpus N, minus the CUDA streams)
        m_local = 2500
        lo = rank * m_local
        og = sharding.OverlappedHitGather(m_local, "cpu", chunks=3)

        def cast(clo, chi, toi_out, tri_out):
            t, k = mesh.cast_rays(None, rays[lo + clo:lo + chi], FMAX)
            toi_out.copy_(torch.from_numpy(t))
            tri_out.copy_(torch.from_numpy(k.astype(np.int32)))
        og.run(cast)
        gt, gk = og.full()
        ok_rays = ok_rays and bool((gt.numpy().view(np.uint32) == rtoi[:world * m_local].view(np.uint32)).all()
                                   and (gk.numpy().view(np.uint32) == rtri[:world * m_local]).all())
        cnt = sharding.all_gather_counts(len(idx), "cpu")
        ok_counts = bool(int(cnt[rank]) == len(idx) and int(cnt.sum()) == len(ridx))
        ret[rank] = (ok_rays, ok_contacts, ok_counts)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from parry_b200 import sharding
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process():
    from harness import oracle
    oracle.build()
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True), (r, ret[r])
