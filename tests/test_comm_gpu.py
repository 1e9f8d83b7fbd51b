"""The C ABI's multi-GPU exchange layer (pb2_comm_*: NCCL through dlopen, CUDA IPC peer buffers) and the sharded broad phase
(pb2_bvh_self_pairs_shard) on two ranks: harness/comm_check.py launched with torch.distributed.run, one process per GPU. Needs two
GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

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
