"""The C ABI's multi-GPU exchange layer (pb2_comm_*: NCCL through dlopen, CUDA IPC peer buffers) and the sharded broad phase
(pb2_bvh_self_pairs_shard) on two ranks: harness/comm_check.py launched with torch.distributed.run, one process per GPU. Needs two
GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_layer_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "harness", "comm_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "comm_check ok" in r.stdout


def test_single_rank_comm_and_shards(ctx):
    """One rank: pb2_comm with nranks = 1 (gathers are copies), and the shards of the self-pair walk partition the pair set."""
    import numpy as np
    import torch
    import parry_b200
    from harness import scenes
    comm = parry_b200.Comm(ctx, parry_b200.Comm.unique_id(), 0, 1)
    assert comm.allgather_counts(42) == [42]
    with torch.cuda.stream(ctx.torch_stream()):
        rows = torch.arange(30, dtype=torch.int32, device="cuda").reshape(10, 3)
    ctx.synchronize()
    out, counts = comm.allgatherv(rows, capacity=4)
    ctx.synchronize()
    assert counts == [10] and (out.cpu().numpy() == rows.cpu().numpy()).all()
    comm.barrier()
    ctx.synchronize()
    n = 20000
    kinds, params, poses, _ = scenes.colliders(n, seed=77)
    shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) for k, p in zip(kinds, params)])
    aabbs = shapes.compute_aabbs(np.arange(n, dtype=np.uint32), poses)
    key = lambda p: np.sort(np.asarray(p).astype(np.int64) @ np.array([1 << 32, 1]))
    for strategy in (0, 1):
        bvh = parry_b200.Bvh.from_leaves(ctx, strategy, aabbs)
        full = key(bvh.traverse_bvtt_single_tree())
        for k in (2, 3, 8):
            parts = [np.asarray(bvh.traverse_bvtt_single_tree_shard(s, k)) for s in range(k)]
            allp = key(np.concatenate(parts))
            assert len(allp) == len(full) and (allp == full).all()
            assert max(len(p) for p in parts) < 1.5 * len(full) / k + 64
    comm.close()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("m,chunks", [(4096, 2), (5000, 3), (100003, 4), (1 << 20, 8), (70000, 16)])
def test_ray_gather_ranges_on_one_gpu(ctx, m, chunks):
    """pb2_trimesh_cast_rays_allgather with a second buffer on the same device standing in for the peer: the range-signalling ray
    kernel (ranges published by per-warp counts while it runs, copy-engine streams waiting on their flags) must give the plain cast's
    results, and the pushed copy must be complete — for batch sizes that are not multiples of the (power-of-two) range size too."""
    import ctypes as C
    import torch
    import parry_b200
    from harness import scenes
    FMAX = float(np.finfo(np.float32).max)
    v, i = scenes.terrain(257, 257)
    mesh = parry_b200.TriMesh(ctx, v, i)
    rays = torch.from_numpy(scenes.terrain_rays(m, seed=21)).cuda()
    toi = torch.zeros(m, dtype=torch.float32, device="cuda")
    tri = torch.zeros(m, dtype=torch.int32, device="cuda")
    ref_toi, ref_tri = mesh.cast_local_ray(rays, FMAX)
    toi2, tri2 = torch.zeros_like(toi), torch.zeros_like(tri)
    p_toi = (C.c_void_p * 2)(toi.data_ptr(), toi2.data_ptr())
    p_tri = (C.c_void_p * 2)(tri.data_ptr(), tri2.data_ptr())
    for _ in range(2):      # the second call finds the flags of the first one reset
        toi.zero_(); tri.zero_(); toi2.zero_(); tri2.zero_()
        torch.cuda.synchronize()
        mesh.cast_local_ray_allgather(rays, FMAX, p_toi, p_tri, 0, 0, chunks)
        ctx.synchronize()
        assert (tri == ref_tri).all() and (toi.view(torch.int32) == ref_toi.view(torch.int32)).all()
        assert (tri2 == ref_tri).all() and (toi2.view(torch.int32) == ref_toi.view(torch.int32)).all()
    assert (ref_tri != -1).float().mean() > 0.3
