#!/usr/bin/env python
"""bench.py — rays/s (TriMesh::cast_ray) and contact pairs/s on B200 vs the host CPU (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input. The headline line (`metric` = rays/s) is
BASELINE config[3]'s per-GPU shard — 2^23 incoherent rays vs the 8,000,000-triangle terrain TriMesh, BVH replicated
per GPU, weak scaling (2^26 rays at 8 GPUs) — because it is the ray configuration whose inputs exceed L2. The other
single-GPU configurations (1M rays vs 1M-triangle sphere, 4M convex pairs, 1M-collider broadphase) are measured in
the same run and reported under "also".

  value : rays/s with rays + outputs resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e   : rays/s through the C ABI with HOST buffers (pinned), H2D + kernels + D2H inside the timed region
  --impl reference : the CPU restatement of parry3d's algorithm (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FMAX = float(np.finfo(np.float32).max)

# Names of the dominant kernels as they are built now. roofline.traffic (ncu dram__bytes_read.sum + dram__bytes_write.sum per
# launch) is read from profiles/r2_traffic.json and only used when the capture there was taken from a kernel of the same name AND
# version string; otherwise it is null (a stale capture must not stand in for the current code).
RAY_KERNEL = "k_raycast_wide_shared<false>"
RAY_KERNEL_VERSION = "r2.3 staged 32-ray refill; 96 B nodes / 64 B triangles read with LDG.E.256; order-independent ties; 8 CTAs per SM"
EPA_KERNEL = "k_contact_epa2"
EPA_KERNEL_VERSION = "r2.2 32-byte face records (one 256-bit load), neighbour records requested with the vertex, 16-entry shared heap head"


def profiled_traffic(kernel, version):
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(p) as f:
            for e in json.load(f):
                if e.get("kernel") == kernel and e.get("kernel_version") == version:
                    return int(e["dram_bytes_per_launch"]), e.get("source")
    except Exception:
        pass
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- workloads
def terrain_scene():
    from harness import scenes
    v, i = scenes.terrain(2001, 2001)
    return v, i


def cpu_rays(v, i, rays, threads, repeat=1):
    """Times the oracle (C++ restatement of TriMesh::cast_ray) on `rays`; returns (rays/s, seconds, build seconds)."""
    from harness import oracle
    t0 = time.perf_counter()
    om = oracle.TriMesh(v, i)
    t_build = time.perf_counter() - t0
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        om.cast_rays(None, rays, FMAX, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(rays) / best, best, t_build, om


def bench_config(args, nt):
    """`config` is the same object in both arms (ours and --impl reference) so that the driver's same_config holds."""
    m = 1 << args.rays_log2
    return {"workload": "2^%d incoherent rays per GPU vs 8,000,000-triangle terrain TriMesh (BASELINE config[3] shard), BVH replicated per GPU"
                        % args.rays_log2,
            "triangles": int(nt), "rays_per_gpu": m, "ray_seed": "6 + rank",
            "l2": "inputs larger than L2 (rays %d MB + results %d MB per step, scene > 600 MB)" % (m * 24 >> 20, m * 8 >> 20)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; parry3d itself cannot be built offline: no cargo,
    nalgebra not vendored) with all host threads. Each step casts the SAME 2^23 rays as one step of our arm on rank 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from harness import oracle, scenes
    oracle.build()
    threads = oracle.hardware_threads()
    v, i = terrain_scene()
    m = 1 << args.rays_log2
    rays = scenes.terrain_rays(m, seed=6)
    om = oracle.TriMesh(v, i)
    args.warmup = max(args.warmup, 1)   # a cold first pass (page faults on the 8M-triangle tree) would understate the reference
    for _ in range(args.warmup):
        om.cast_rays(None, rays, FMAX, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        om.cast_rays(None, rays, FMAX, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = m / dt
    line = {
        "impl": "reference", "metric": "rays/s (TriMesh::cast_ray)", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, len(i)),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": "all %d rays of rank 0's shard per step, all host threads (C++ restatement of parry3d's TriMesh::cast_ray)" % m},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def oracle_slice_check(om, oracle, rays_h, toi, tri, k):
    """Checks the first k rays of the TIMED device output against the CPU oracle (bit-exact toi + id; the two order-dependent cases
    of DESIGN.md section 3 are adjudicated by brute force over all triangles). Raises on any unexplained difference."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import check_ray_parity
    r = om.cast_rays(None, rays_h[:k], FMAX, threads=oracle.hardware_threads())

    def brute(idx):
        t, i, _, _ = om.cast_rays(None, rays_h[:k][idx], FMAX, with_normal=True, mode=1, threads=oracle.hardware_threads())
        return t, i
    g = (toi[:k], tri[:k])
    adjudicated = check_ray_parity(g, r, brute, max_ulp_cases=1e-4)
    same = int(((np.asarray(g[1]).astype(np.uint32) == np.asarray(r[1]).astype(np.uint32)) &
                (np.asarray(g[0]).view(np.uint32) == np.asarray(r[0]).view(np.uint32))).sum())
    return {"checker": "CPU oracle (reference-order traversal) on the first %d rays of the timed output" % k, "rays": k, "bit_exact": same,
            "adjudicated_by_brute_force": int(adjudicated), "hits": int((np.asarray(r[1]) != 0xFFFFFFFF).sum())}


def cpu_side_baselines(oracle, threads):
    """CPU (oracle port) figures for the other two named workloads, on bounded samples: query::contact on hull pairs of the 4M-pair
    scene (1 thread and all threads) and one broad-phase frame on 2^20 colliders (sequential in the reference: 1 thread)."""
    from harness import scenes
    out = {}
    pts, radii = scenes.hull_pool(4096)
    T = oracle.ShapeTable([("convex", p) for p in pts])
    n_all, n_one = 1 << 18, 1 << 15
    a, b, p1, p2 = scenes.hull_pairs(n_all, radii, seed=4)
    t0 = time.perf_counter()
    T.contact(a, p1, b, p2, 0.01, threads=threads)
    dt_all = time.perf_counter() - t0
    t0 = time.perf_counter()
    T.contact(a[:n_one], p1[:n_one], b[:n_one], p2[:n_one], 0.01, threads=1)
    dt_one = time.perf_counter() - t0
    out["contacts"] = {"value": n_all / dt_all, "unit": "pairs/s", "cores": threads, "one_thread_value": n_one / dt_one, "kind": "port",
                       "sample": "first %d (all threads) / %d (1 thread) pairs of the 2^22-pair scene (seed 4), 32-vertex hulls, prediction 0.01"
                                 % (n_all, n_one)}
    n = 1 << 20
    kinds, params, poses, _ = scenes.colliders(n, seed=2)
    t0 = time.perf_counter()
    aabbs = oracle.shape_aabbs(kinds, params, poses)
    ob = oracle.Bvh(aabbs)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    aabbs = oracle.shape_aabbs(kinds, params, poses)
    ob.update_leaves(aabbs, np.arange(n, dtype=np.uint32), 0.0)
    ob.refit()
    pr = ob.self_pairs()
    dt = time.perf_counter() - t0
    out["broadphase"] = {"value": n / dt, "unit": "AABBs/s (frame: leaf AABBs + update + refit + traverse_bvtt_single_tree)", "cores": 1, "kind": "port",
                         "pairs_per_frame": int(len(pr)), "pairs_per_s": len(pr) / dt, "frame_ms": dt * 1e3,
                         "build_from_leaves_ms": t_build * 1e3,
                         "sample": "one frame on the full 2^20-collider scene (seed 2); the reference's Bvh is sequential (bvh_tree.rs:24)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-log2", type=int, default=23, help="rays per GPU = 2^k")
    ap.add_argument("--skip-also", action="store_true", help="only the headline workload")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--first-hw-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.first_hw_child:
        first_hw_child()
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import parry_b200
    from harness import scenes

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = parry_b200.Context(local_rank)
    stream = ctx.torch_stream()
    hbm_peak, peak_src = load_peaks()

    # ------------------------------------------------------------------ headline: terrain rays
    v, i = terrain_scene()
    mesh = parry_b200.TriMesh(ctx, v, i)
    m = 1 << args.rays_log2
    rays_h = scenes.terrain_rays(m, seed=6 + rank)  # each rank casts its own shard of the 2^26-ray set
    rays_d = torch.from_numpy(rays_h).cuda()
    toi_d = torch.empty(m, dtype=torch.float32, device="cuda")
    tri_d = torch.empty(m, dtype=torch.int32, device="cuda")
    gather, gather_kind, peer = None, "none (N=1)", False
    # the C ABI's own exchange layer (pb2_comm_*: NCCL + CUDA IPC peer buffers); torch.distributed only carries the 128-byte id
    comm = parry_b200.Comm.from_torch_distributed(ctx) if world > 1 else parry_b200.Comm(ctx, parry_b200.Comm.unique_id(), 0, 1)
    if world > 1:
        from parry_b200 import sharding
        chunks = int(os.environ.get("PB2_BENCH_GATHER_CHUNKS", "4"))
        try:
            if os.environ.get("PB2_BENCH_NCCL_GATHER") or os.environ.get("PB2_BENCH_SYMM_GATHER"):
                raise RuntimeError("forced")
            gather = sharding.CommHitGather(comm, m, torch.device("cuda", local_rank), chunks=chunks)
            gather_kind = ("all-gather of the (toi,id) results inside the timed region through the C ABI: finished result ranges are pushed into every "
                           "peer's buffer (pb2_comm_peer_alloc: CUDA IPC) by copy engines over NVLink while the traversal kernel runs, then pb2_comm_barrier")
        except Exception as e0:
            try:
                if os.environ.get("PB2_BENCH_NCCL_GATHER"):
                    raise RuntimeError("forced")
                gather = sharding.PeerHitGather(m, torch.device("cuda", local_rank), chunks=chunks)
                gather_kind = ("all-gather of the (toi,id) results inside the timed region: pieces pushed into every peer's symmetric-memory "
                               "buffer by copy engines over NVLink while the next piece is traversed, then a cross-rank barrier (%s)" % type(e0).__name__)
            except Exception as e:  # no peer mappings on this box: NCCL all_gather per piece on a side stream
                gather = sharding.OverlappedHitGather(m, "cuda", chunks=4)
                gather_kind = "all_gather (NCCL) of the (toi,id) results inside the timed region, 4 pieces on a side stream (%s)" % type(e).__name__
        peer = isinstance(gather, (sharding.PeerHitGather, sharding.CommHitGather))

    def step_device():
        if world == 1:
            mesh.cast_local_ray(rays_d, FMAX, out=(toi_d, tri_d))
        elif peer:
            gather.run(mesh, rays_d, FMAX, stream)
        else:
            gather.run(lambda lo, hi, t, k: mesh.cast_local_ray(rays_d[lo:hi], FMAX, out=(t, k)), stream)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, ctx.launch_count - l0

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.finish()
    value = world * m / (ms_dev * 1e-3)

    # kernel-only duration for the roofline (same launches, events on the launching stream, no collective)
    def step_kernel():
        mesh.cast_local_ray(rays_d, FMAX, out=(toi_d, tri_d))
    ms_kernel, _ = timed(step_kernel, args.steps, 1)
    toi_timed = toi_d.cpu().numpy().copy()          # what the timed kernel launches wrote (checked against the oracle below)
    tri_timed = tri_d.cpu().numpy().view(np.uint32).copy()

    # end-to-end: host (pinned) buffers through the C ABI, copies inside the timed region
    rays_pin = torch.from_numpy(rays_h).pin_memory()
    toi_pin = torch.empty(m, dtype=torch.float32).pin_memory()
    tri_pin = torch.empty(m, dtype=torch.int32).pin_memory()
    rays_np, toi_np, tri_np = rays_pin.numpy(), toi_pin.numpy(), tri_pin.numpy().view(np.uint32)

    def step_e2e():
        mesh.cast_local_ray(rays_np, FMAX, out=(toi_np, tri_np))

    for _ in range(2):
        step_e2e()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        step_e2e()
    dt_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([dt_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_e2e = float(t.item())
    e2e_val = world * m / dt_e2e
    # the host path returns the same bits as the device path
    paths_differ = int(((toi_timed.view(np.uint32) != toi_np.view(np.uint32)) | (tri_timed != tri_np)).sum())
    assert paths_differ == 0, "device-resident and host-buffer ray casts differ on %d rays" % paths_differ
    if world > 1:
        # every rank must hold every rank's results, element by element: each rank checks the shard of its right-hand neighbour
        # against that neighbour's own local results (toi bits and triangle ids), plus per-shard checksums of all shards
        ft, fk = gather.full()
        torch.cuda.synchronize()
        mine = torch.tensor([int(fk[rank * m:(rank + 1) * m].to(torch.int64).sum().item())], device="cuda")
        owners = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(owners, mine)
        seen = torch.stack([fk[r * m:(r + 1) * m].to(torch.int64).sum() for r in range(world)])
        assert bool((seen == owners).all()), "gathered results differ from the owners' results"
        nb = (rank + 1) % world
        loc_t, loc_k = torch.from_numpy(toi_timed).cuda(), torch.from_numpy(tri_timed.view(np.int32)).cuda()
        all_t = [torch.empty_like(loc_t) for _ in range(world)]
        all_k = [torch.empty_like(loc_k) for _ in range(world)]
        dist.all_gather(all_t, loc_t)
        dist.all_gather(all_k, loc_k)
        assert bool((ft[nb * m:(nb + 1) * m].view(torch.int32) == all_t[nb].view(torch.int32)).all()), "gathered toi differs from the owner's"
        assert bool((fk[nb * m:(nb + 1) * m].view(torch.int32) == all_k[nb]).all()), "gathered triangle ids differ from the owner's"
        del all_t, all_k, loc_t, loc_k

    nt, nv = len(i), len(v)
    # Algorithmic (compulsory) bytes per launch. Contract figure: 32 B per ray (24 in + 8 out) + the scene read once in
    # the reference's own layout: 64 (T - 1) node bytes + 12 V vertex bytes + 12 T index bytes. Also reported: the same with this
    # library's scene layout (96-byte wide nodes + 64-byte pre-gathered triangles), which is what the kernel actually has to touch.
    alg_contract = m * 32 + 64 * (nt - 1) + 12 * nv + 12 * nt
    alg_own = m * 32 + mesh.wide_bytes() if hasattr(mesh, "wide_bytes") else None
    achieved = alg_contract / (ms_kernel * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(RAY_KERNEL, RAY_KERNEL_VERSION)
    line = {
        "metric": "rays/s (TriMesh::cast_ray)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, nt),
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": m * 24, "d2h_bytes_per_step": m * 8,
                "ms_per_step": dt_e2e * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": RAY_KERNEL, "kernel_version": RAY_KERNEL_VERSION,
                     "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_contract,
                     "algorithmic_bytes_rule": "SURVEY 8(d): 32 B/ray x 2^%d + 64(T-1) + 12V + 12T (the reference's own scene layout)" % args.rays_log2,
                     "peak_source": peak_src, "collective": gather_kind},
    }
    if alg_own:
        line["roofline"]["own_layout"] = {"algorithmic_bytes_per_launch": alg_own, "achieved": alg_own / (ms_kernel * 1e-3) / 1e9,
                                          "frac": alg_own / (ms_kernel * 1e-3) / 1e9 / hbm_peak}

    # ------------------------------------------------------------------ the second named metric: contact pairs/s (BASELINE config[2])
    timed_also = make_timed(ctx, stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if not args.skip_also else None   # > 126 MB L2
    also = {}
    if not args.skip_also:
        r = also_contacts(ctx, stream, timed_also, None, hbm_peak, seed=4 + rank, e2e=(world == 1))
        ms_c = r["ms"]
        if world > 1:   # every rank its own 2^22-pair shard (weak scaling, no data-path collective), device time = max over ranks
            t = torch.tensor([ms_c], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_c = float(t.item())
        pairs_s = world * r["pairs"] / (ms_c * 1e-3)
        line["roofline"]["second_metric"] = {
            "metric": "contact pairs/s (query::contact, 2^22 ConvexPolyhedron pairs per GPU, 32-vertex hulls, prediction 0.01; BASELINE config[2])",
            "value": pairs_s, "unit": "pairs/s", "ms_per_step": ms_c, "n_gpus": world, "scaling": "weak",
            "e2e_value": r.get("e2e_value"), "contacts_fraction": r["contacts_fraction"],
            "kernel": EPA_KERNEL, "kernel_version": EPA_KERNEL_VERSION, "kernel_ms": r.get("epa_ms"), "gjk_kernel_ms": r.get("gjk_ms"),
            "epa_runs": r.get("epa_runs"), "algorithmic_bytes_per_launch": r["pairs"] * 120,
            "algorithmic_bytes_rule": "SURVEY 8(d): 120 B/pair (64 in + 56 out), hull pool L2-resident",
            "achieved": r["pairs"] * 120 / (ms_c * 1e-3) / 1e9, "frac": r["pairs"] * 120 / (ms_c * 1e-3) / 1e9 / hbm_peak,
            "dominant_kernel_frac": (r["pairs"] * 120 / (r["epa_ms"] * 1e-3) / 1e9 / hbm_peak) if r.get("epa_ms") else None,
            "traffic": profiled_traffic(EPA_KERNEL, EPA_KERNEL_VERSION)[0],
            "shallow_mix_value": (r.get("shallow_mix") or {}).get("value")}
        also["contact_pairs_4M_hulls"] = r
        for name, fn in MULTI_GPU_ALSO:   # collective: every rank runs these, at every N (N = 1 gives the series its first point)
            try:
                also[name] = fn(ctx, stream, timed_also, flush, hbm_peak, comm, dist, rank, world)
            except Exception as e:
                also[name] = {"error": repr(e)}
            # BASELINE config[4] at N GPUs, where the driver keeps it (N2 had no kept record)
            line["roofline"].setdefault("multi_gpu", {})[name] = {k: also[name].get(k) for k in (
                "value", "unit", "ms", "n_gpus", "scaling", "pairs_per_frame", "contacts_per_frame", "pairs_per_s", "gathered_bytes_per_rank",
                "all_ranks_agree", "error") if k in also[name]}
        if rank == 0 and world == 1:
            # single-GPU extras (spheres, manifolds, sibling queries, broad phase, TriMesh contacts): measured at N = 1 only, so that
            # at N > 1 the other ranks do not sit in the closing barrier while rank 0 works through them
            also.update(bench_also(ctx, stream, args, hbm_peak, flush))

    if rank == 0 and not args.skip_cpu:
        from harness import oracle
        oracle.build()
        threads = oracle.hardware_threads()
        val, secs, t_build, om = cpu_rays(v, i, rays_h, threads)
        one = 1 << 18
        t0 = time.perf_counter()
        om.cast_rays(None, rays_h[:one], FMAX, threads=1)
        dt_one = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": val, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": "all %d rays of rank 0's shard, same mesh, all host threads: %.2f s cast (+ %.2f s Bvh build, 1 thread); "
                                          "C++ restatement of parry3d (no Rust toolchain on the box)" % (m, secs, t_build),
                                "one_thread_value": one / dt_one, "one_thread_sample": "first %d rays" % one}
        # the timed device output against the oracle (not against itself)
        line["parity"] = oracle_slice_check(om, oracle, rays_h, toi_timed, tri_timed, 1 << 16)
        del om
        if not args.skip_also:
            try:
                line["cpu_baseline"].update(cpu_side_baselines(oracle, threads))
            except Exception as e:
                line["cpu_baseline"]["also_error"] = repr(e)
    if also and rank == 0:
        # details of the secondary workloads go to a side file; the line keeps one number per workload so that it stays readable
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "bench_also_n%d.json" % world), "w") as f:
                json.dump(also, f, indent=1)
        except Exception:
            pass
        def brief(val):
            if not isinstance(val, dict):
                return val
            if "value" in val or "error" in val:
                return {kk: vv for kk, vv in val.items() if kk in ("value", "unit", "ms", "error", "n_gpus")}
            return {kk: brief(vv) for kk, vv in val.items() if isinstance(vv, dict)}   # groups of workloads: one number each
        line["also"] = {k: brief(val) for k, val in also.items()}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def make_timed(ctx, stream):
    import torch

    def timed(fn, steps=10, warmup=3, flush=None):
        for _ in range(warmup):
            fn()
        ctx.synchronize()
        tot = 0.0
        for _ in range(steps):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / steps
    return timed


def bench_also(ctx, stream, args, hbm_peak, flush=None):
    """Secondary single-GPU configurations, device-timed the same way (inputs resident, CUDA events)."""
    import torch
    import parry_b200
    from harness import scenes
    out = {}

    def timed(fn, steps=10, warmup=3, flush=None):
        for _ in range(warmup):
            fn()
        ctx.synchronize()
        tot = 0.0
        for _ in range(steps):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / steps

    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # 1M rays vs 1M-triangle sphere (north_star target configuration); inputs fit in L2 => flush between iterations
    v, i = scenes.uv_sphere(708, 707)
    mesh = parry_b200.TriMesh(ctx, v, i)
    m = 1 << 20
    rays = torch.from_numpy(scenes.sphere_rays(m, seed=1)).cuda()
    toi = torch.empty(m, dtype=torch.float32, device="cuda")
    tri = torch.empty(m, dtype=torch.int32, device="cuda")
    ms = timed(lambda: mesh.cast_local_ray(rays, FMAX, out=(toi, tri)), flush=flush)
    alg = m * 32 + 64 * (len(i) - 1) + 12 * len(v) + 12 * len(i)
    out["rays_1M_vs_1M_tri_sphere"] = {"value": m / (ms * 1e-3), "unit": "rays/s", "ms": ms, "l2": "flushed between iterations",
                                       "roofline_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
    v, i = scenes.uv_sphere(224, 224)
    mesh2 = parry_b200.TriMesh(ctx, v, i)
    ms = timed(lambda: mesh2.cast_local_ray(rays, FMAX, out=(toi, tri)), flush=flush)
    alg = m * 32 + 64 * (len(i) - 1) + 12 * len(v) + 12 * len(i)
    out["rays_1M_vs_100k_tri_sphere"] = {"value": m / (ms * 1e-3), "unit": "rays/s", "ms": ms, "l2": "flushed between iterations",
                                         "roofline_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
    del mesh, mesh2
    for name, fn in EXTRA_ALSO:
        try:
            out[name] = fn(ctx, stream, timed, flush, hbm_peak)
        except Exception as e:  # a secondary workload must not take the headline down
            out[name] = {"error": repr(e)}
    if int(os.environ.get("WORLD_SIZE", "1")) == 1:
        # manifold persistence and Compound-vs-Compound: in a child process with a timeout (the two largest allocations of the run), so that
        # neither a device fault nor a crash or hang there can reach this process and its JSON line
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(ctx.device)))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--first-hw-child"], capture_output=True, text=True, timeout=420, env=env)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out["persistence_and_compound_pairs"] = json.loads(lines[-1]) if lines else {"error": "child rc=%d: %s" % (r.returncode, r.stderr[-300:])}
        except Exception as e:
            out["persistence_and_compound_pairs"] = {"error": repr(e)}
    return out


def first_hw_child():
    """Child process of bench_also: times manifold persistence and Compound-vs-Compound contacts on device 0 of its own CUDA context and prints one JSON object."""
    try:
        import torch
        import parry_b200
        hbm_peak, _ = load_peaks()
        torch.cuda.set_device(0)
        ctx = parry_b200.Context(0)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        res = also_persistence_and_compounds(ctx, ctx.torch_stream(), make_timed(ctx, ctx.torch_stream()), flush, hbm_peak)
    except Exception as e:
        res = {"error": repr(e)}
    print(json.dumps(res), flush=True)


def also_contacts(ctx, stream, timed, flush, hbm_peak, seed=4, e2e=True):
    """BASELINE config[2]: 2^22 ConvexPolyhedron pairs (32-vertex hulls from a pool of 4096), prediction 0.01."""
    import torch
    import parry_b200
    from harness import scenes
    pts, radii = scenes.hull_pool(4096)
    G = parry_b200.Shapes(ctx, [parry_b200.ConvexPolyhedron(p) for p in pts])
    n = 1 << 22
    a, b, p1, p2 = scenes.hull_pairs(n, radii, seed=seed)
    da, db = torch.from_numpy(a.astype(np.int32)).cuda(), torch.from_numpy(b.astype(np.int32)).cuda()
    dp1, dp2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    res = {}

    def run():
        res["out"] = parry_b200.contact(G, da, dp1, db, dp2, 0.01)
    ms = timed(run, steps=5, warmup=3)  # inputs+outputs = 2^22 * 120 B = 503 MB > L2
    phase = None
    try:   # per-kernel durations of one more call (CUDA events on the library's stream around the GJK / EPA / finishing kernels)
        ctx.enable_phase_timing(True)
        run()
        phase = ctx.contact_phase_times()
        ctx.enable_phase_timing(False)
    except Exception:
        pass
    st = res["out"][1]
    frac_some = float((st == 1).float().mean().item())
    alg = n * 120
    r = {"value": n / (ms * 1e-3), "unit": "pairs/s", "ms": ms, "pairs": n, "contacts_fraction": frac_some,
         "l2": "inputs+outputs (503 MB) larger than L2", "roofline_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak,
         "algorithmic_bytes_per_pair": 120}
    if phase:
        r.update({"gjk_ms": phase[0], "epa_ms": phase[1], "finish_ms": phase[2], "epa_runs": phase[3]})
    if not e2e:
        return r
    # same batch size with the separations of a settled scene (18 % of the pairs penetrate instead of 67 %): GJK-only
    # pairs cost ~1.2 ns, EPA runs ~9 ns, so the mix decides the pairs/s figure
    a2, b2, q1, q2 = scenes.hull_pairs(n, radii, seed=seed + 1, s_lo=1.7, s_hi=2.4)
    ea, eb = torch.from_numpy(a2.astype(np.int32)).cuda(), torch.from_numpy(b2.astype(np.int32)).cuda()
    eq1, eq2 = torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda()

    def run_shallow():
        res["shallow"] = parry_b200.contact(G, ea, eq1, eb, eq2, 0.01)
    ms2 = timed(run_shallow, steps=5, warmup=3)
    o2, st2 = res["shallow"]
    r["shallow_mix"] = {"value": n / (ms2 * 1e-3), "unit": "pairs/s", "ms": ms2, "separation": "[1.7, 2.4] x mean radius",
                        "contacts_fraction": float((st2 == 1).float().mean().item()),
                        "penetrating_fraction": float(((st2 == 1) & (o2[:, 12] < 0)).float().mean().item())}
    del ea, eb, eq1, eq2, o2, st2
    res.pop("shallow")
    # end to end through the C ABI with pinned host buffers (H2D + kernels + D2H inside the timed region)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
    ha, hb, hp1, hp2 = pin(a.astype(np.uint32).view(np.int32)).view(np.uint32), pin(b.astype(np.uint32).view(np.int32)).view(np.uint32), pin(p1), pin(p2)
    hout = (torch.empty((n, 13), dtype=torch.float32).pin_memory().numpy(), torch.empty(n, dtype=torch.uint8).pin_memory().numpy())
    parry_b200.contact(G, ha, hp1, hb, hp2, 0.01, out=hout)
    t0 = time.perf_counter()
    for _ in range(3):
        parry_b200.contact(G, ha, hp1, hb, hp2, 0.01, out=hout)
    r["e2e_value"] = n / ((time.perf_counter() - t0) / 3)
    return r


def also_broadphase(ctx, stream, timed, flush, hbm_peak):
    """BASELINE config[1]: 2^20 dynamic colliders (balls + cuboids): per frame AABBs -> Bvh (rebuild, or update + refit)
    -> full self-pair set."""
    import torch
    import parry_b200
    from harness import scenes
    n = 1 << 20
    kinds, params, poses, _ = scenes.colliders(n, seed=2)
    shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) for k, p in zip(kinds, params)])
    ids = torch.arange(n, dtype=torch.int32, device="cuda")
    dposes = torch.from_numpy(poses).cuda()
    aabbs = shapes.compute_aabbs(ids, dposes)
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    cap = 16 * n
    state = {}

    def frame_rebuild():
        a = shapes.compute_aabbs(ids, dposes)
        bvh.insert_or_update_partially(a, ids, 0.0)
        bvh.rebuild()
        state["pairs"] = bvh.traverse_bvtt_single_tree(capacity=cap, like=a)

    def frame_refit():
        a = shapes.compute_aabbs(ids, dposes)
        bvh.insert_or_update_partially(a, ids, 0.0)
        bvh.refit()
        state["pairs"] = bvh.traverse_bvtt_single_tree(capacity=cap, like=a)

    ms_rebuild = timed(frame_rebuild, steps=5, warmup=3, flush=flush)
    npairs = int(state["pairs"].shape[0])
    ms_refit = timed(frame_refit, steps=5, warmup=3, flush=flush)
    alg = 24 * n + 8 * npairs
    return {"value": n / (ms_rebuild * 1e-3), "unit": "AABBs/s (frame: AABBs + rebuild + self pairs)", "ms": ms_rebuild,
            "pairs_per_frame": npairs, "pairs_per_s": npairs / (ms_rebuild * 1e-3),
            "refit_frame_ms": ms_refit, "refit_frame_aabbs_per_s": n / (ms_refit * 1e-3), "l2": "flushed between iterations",
            "roofline_frac": alg / (ms_rebuild * 1e-3) / 1e9 / hbm_peak}


def also_mixed(ctx, stream, timed, flush, hbm_peak):
    """BASELINE config[4]: 2^21 ball / cuboid / 32-vertex-hull colliders; per frame leaf AABBs -> Bvh build -> self pair
    set -> query::contact on every pair (compacted contacts), all device resident."""
    import torch
    import parry_b200
    from harness import scenes
    n, H = 1 << 21, 4096
    kinds, params, poses, hull_ids = scenes.colliders(n, seed=8, hull_fraction=1.0 / 3.0, n_hulls=H)
    pts, _ = scenes.hull_pool(H, 32, seed=9)
    pts = np.asarray(pts, dtype=np.float32) * 0.5
    tk = np.concatenate([np.full(H, 2, np.uint8), np.where(kinds == 2, 0, kinds).astype(np.uint8)])
    tp = np.concatenate([np.zeros((H, 3), np.float32), params])
    first = np.concatenate([np.arange(H, dtype=np.uint32) * 32, np.zeros(n, np.uint32)])
    count = np.concatenate([np.full(H, 32, np.uint32), np.zeros(n, np.uint32)])
    G = parry_b200.Shapes.from_arrays(ctx, tk, tp, pts.reshape(-1, 3), first, count)
    cs = torch.from_numpy(np.where(kinds == 2, hull_ids, H + np.arange(n)).astype(np.uint32).view(np.int32)).cuda()
    dposes = torch.from_numpy(poses).cuda()
    aabbs = G.compute_aabbs(cs, dposes)
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    state = {}

    def frame():
        a = G.compute_aabbs(cs, dposes)
        bvh.insert_or_update_partially(a, torch.arange(n, dtype=torch.int32, device="cuda"), 0.0)
        bvh.rebuild()
        pr = bvh.traverse_bvtt_single_tree(capacity=16 * n, like=a)
        state["pairs"] = int(pr.shape[0])
        out, idx = parry_b200.contact_pairs_compact(G, cs, dposes, pr, 0.01)
        state["contacts"] = int(out.shape[0])

    ms = timed(frame, steps=5, warmup=2, flush=flush)
    P, Cn = state["pairs"], state["contacts"]
    alg = 24 * n + 8 * P + P * 8 + Cn * 56
    out = {"value": n / (ms * 1e-3), "unit": "colliders/s (frame: AABBs + Bvh build + self pairs + contacts)", "ms": ms,
           "colliders": n, "pairs_per_frame": P, "contacts_per_frame": Cn, "pairs_per_s": P / (ms * 1e-3),
           "l2": "flushed between iterations", "roofline_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
    # the same frame ending in per-pair contact MANIFOLDS (BASELINE config[4] as worded): the pair list is expanded to per-pair
    # shape / pose arrays on the device, then pb2_contact_manifolds_batch (all arms; first-frame manifolds). Hull topology for the
    # pool comes from the test-side restatement of ConvexPolyhedron::from_convex_mesh, so this variant uses the first 256 pool hulls.
    try:
        from harness import hull_topology as ht
        H2 = 256
        hulls = [pts[h] for h in range(H2)]
        t = ht.hull_table(hulls)
        nsh = H + n
        hf, hc, hef = np.zeros(nsh, np.uint32), np.zeros(nsh, np.uint32), np.zeros(nsh, np.uint32)
        hf[:H2], hc[:H2], hef[:H2] = t["hull_face_first"], t["hull_face_count"], t["hull_edge_first"]
        hf[H2:H], hc[H2:H] = t["hull_face_first"][0], t["hull_face_count"][0]     # unused pool entries: any valid range
        npnt = H * 32
        vf, vc = np.zeros(npnt, np.uint32), np.zeros(npnt, np.uint32)
        vf[:H2 * 32], vc[:H2 * 32] = t["vert_first"], t["vert_count"]
        G.set_hull_topology(dict(t, hull_face_first=hf, hull_face_count=hc, hull_edge_first=hef, vert_first=vf, vert_count=vc))
        cs2 = torch.where(cs < H, cs % H2, cs)     # hull colliders draw from the 256 hulls that have topology
        torch.cuda.synchronize()

        def frame_manifolds():
            # the gathers are torch kernels: they must run on the library's stream, or the library would read the per-pair arrays
            # before torch has written them
            with torch.cuda.stream(stream):
                a = G.compute_aabbs(cs2, dposes)
                bvh.insert_or_update_partially(a, torch.arange(n, dtype=torch.int32, device="cuda"), 0.0)
                bvh.rebuild()
                pr = bvh.traverse_bvtt_single_tree(capacity=16 * n, like=a).to(torch.int64)
                i1, i2 = pr[:, 0], pr[:, 1]
                state["m"] = parry_b200.contact_manifolds(G, cs2[i1], dposes[i1], cs2[i2], dposes[i2], 0.01, max_points=10)
        ms2 = timed(frame_manifolds, steps=3, warmup=2, flush=flush)
        cnt, st = state["m"][1], state["m"][3]
        out["with_manifolds"] = {"ms": ms2, "value": n / (ms2 * 1e-3), "pairs_per_s": int(cnt.shape[0]) / (ms2 * 1e-3), "pairs_per_frame": int(cnt.shape[0]),
                                 "manifolds_with_points": int((cnt > 0).sum().item()), "points": int(cnt.to(torch.int64).sum().item()),
                                 "status_counts": torch.bincount(st.to(torch.int64), minlength=5).tolist()}
    except Exception as e:
        out["with_manifolds"] = {"error": repr(e)}
    return out


def also_mesh_contacts(ctx, stream, timed, flush, hbm_peak):
    """2^20 ball / cuboid / hull colliders resting on or near the 8,000,000-triangle terrain TriMesh:
    query::contact(mesh, collider) for every collider (mesh Bvh query + per-triangle narrow phase + min reduction)."""
    import torch
    import parry_b200
    from harness import scenes
    v, i = terrain_scene()
    mesh = parry_b200.TriMesh(ctx, v, i)
    n, H = 1 << 20, 4096
    g = scenes.rng(11)
    pts, _ = scenes.hull_pool(H, 32, seed=12)
    pts = np.asarray(pts, dtype=np.float32) * 0.4
    kinds = g.integers(0, 3, n).astype(np.uint8)
    params = (g.random((n, 3)) * 0.3 + 0.15).astype(np.float32)
    tk = np.concatenate([np.full(H, 2, np.uint8), np.where(kinds == 2, 0, kinds).astype(np.uint8)])
    tp = np.concatenate([np.zeros((H, 3), np.float32), params])
    first = np.concatenate([np.arange(H, dtype=np.uint32) * 32, np.zeros(n, np.uint32)])
    count = np.concatenate([np.full(H, 32, np.uint32), np.zeros(n, np.uint32)])
    G = parry_b200.Shapes.from_arrays(ctx, tk, tp, pts.reshape(-1, 3), first, count)
    sid = np.where(kinds == 2, g.integers(0, H, n), H + np.arange(n)).astype(np.uint32)
    anchor = v[g.integers(0, len(v), n)]
    t = anchor + np.stack([g.standard_normal(n) * 0.2, (g.random(n) - 0.3) * 0.8, g.standard_normal(n) * 0.2], axis=1)
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    dsid, dposes = torch.from_numpy(sid.view(np.int32)).cuda(), torch.from_numpy(poses).cuda()
    mpose = torch.tensor([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float32, device="cuda")
    res = {}

    def run():
        res["o"] = mesh.contact_shapes(mpose, G, dsid, dposes, 0.02)
    ms = timed(run, steps=5, warmup=2, flush=flush)
    st = res["o"][1]
    return {"value": n / (ms * 1e-3), "unit": "colliders/s (TriMesh-vs-shape contacts, 8M-triangle terrain)", "ms": ms, "colliders": n,
            "contacts_fraction": float((st == 1).float().mean().item()), "l2": "flushed between iterations"}


def also_manifolds(ctx, stream, timed, flush, hbm_peak):
    """SURVEY §8 f1 (closed-form arms): 2^22 ball / cuboid pairs -> contact manifolds (ball-ball, ball-cuboid, cuboid-cuboid SAT +
    face clipping), first frame, up to 8 points per manifold."""
    import torch
    import parry_b200
    from harness import scenes
    n = 1 << 22
    g = scenes.rng(21)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(0.4), parry_b200.Ball(0.25), parry_b200.Cuboid([0.3, 0.5, 0.4]), parry_b200.Cuboid([0.6, 0.2, 0.2]),
                                parry_b200.Cuboid([0.5, 0.5, 0.5])])
    s1, s2 = g.integers(0, 5, n).astype(np.int32), g.integers(0, 5, n).astype(np.int32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 20], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.3 + 0.2)], axis=1).astype(np.float32)
    ds1, ds2, dp1, dp2 = (torch.from_numpy(x).cuda() for x in (s1, s2, p1, p2))
    res = {}

    def run():
        res["o"] = parry_b200.contact_manifolds(G, ds1, dp1, ds2, dp2, 0.05, max_points=8)
    ms = timed(run, steps=5, warmup=2, flush=flush)
    cnt = res["o"][1]
    npts = int(cnt.to(torch.int64).sum().item())
    alg = n * (8 + 56 + 24 + 4 + 1) + npts * 36
    out = {"value": n / (ms * 1e-3), "unit": "pairs/s (ball / cuboid contact manifolds)", "ms": ms, "pairs": n,
           "manifolds_with_points": float((cnt > 0).float().mean().item()), "points": npts, "l2": "flushed between iterations",
           "roofline_frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
    # pfm_pfm arm: 2^20 pairs of 32-vertex hulls / cuboids (hull topology built on the host by the test-side restatement of
    # ConvexPolyhedron::from_convex_mesh; parry would hand over its own)
    from harness import hull_topology as ht
    pts, _ = scenes.hull_pool(64, 32, seed=22)
    hulls = [np.asarray(p, np.float32) * 0.6 for p in pts]
    G2 = parry_b200.Shapes(ctx, [parry_b200.Cuboid([0.3, 0.5, 0.4]), parry_b200.Cuboid([0.6, 0.2, 0.2])] + [parry_b200.ConvexPolyhedron(h) for h in hulls])
    t = ht.hull_table(hulls)
    hf, hc, hef = np.zeros(66, np.uint32), np.zeros(66, np.uint32), np.zeros(66, np.uint32)
    hf[2:], hc[2:], hef[2:] = t["hull_face_first"], t["hull_face_count"], t["hull_edge_first"]
    # per table entry; the vertex-side arrays are per point, and the table's points are exactly the hulls' points in order
    t = dict(t, hull_face_first=hf, hull_face_count=hc, hull_edge_first=hef)
    G2.set_hull_topology(t)
    m = 1 << 20
    h1, h2 = g.integers(0, 66, m).astype(np.int32), g.integers(2, 66, m).astype(np.int32)
    dh1, dh2 = torch.from_numpy(h1).cuda(), torch.from_numpy(h2).cuda()

    def run_pfm():
        res["p"] = parry_b200.contact_manifolds(G2, dh1, dp1[:m], dh2, dp2[:m], 0.05, max_points=12)
    ms2 = timed(run_pfm, steps=5, warmup=2, flush=flush)
    c2, st2 = res["p"][1], res["p"][3]
    out["pfm_pfm_1M_hull_pairs"] = {"value": m / (ms2 * 1e-3), "unit": "pairs/s", "ms": ms2, "manifolds_with_points": float((c2 > 0).float().mean().item()),
                                    "host_fallback": int((st2 == 3).sum().item())}
    return out


def _mixed_table(ctx, n_hulls=64, seed=31):
    import parry_b200
    from harness import scenes
    pts, _ = scenes.hull_pool(n_hulls, 32, seed=seed)
    spec = [parry_b200.Ball(0.4), parry_b200.Ball(0.25), parry_b200.Cuboid([0.3, 0.5, 0.4]), parry_b200.Cuboid([0.6, 0.2, 0.2])]
    spec += [parry_b200.ConvexPolyhedron(np.asarray(p, np.float32) * 0.6) for p in pts]
    return parry_b200.Shapes(ctx, spec), len(spec)


def _mixed_pairs(n, ns, seed, spread=2.5):
    from harness import scenes
    g = scenes.rng(seed)
    s1, s2 = g.integers(0, ns, n).astype(np.int32), g.integers(0, ns, n).astype(np.int32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 20], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * spread + 0.2)], axis=1).astype(np.float32)
    return g, d, s1, p1, s2, p2


def also_siblings(ctx, stream, timed, flush, hbm_peak):
    """SURVEY §8 f2/f3 queries on 2^21 mixed ball / cuboid / 32-vertex-hull pairs, device resident: cast_shapes (default
    options, a quarter of the pairs start overlapping), distance, closest_points; compound contacts (2^20 compounds of 1-5 parts)."""
    import torch
    import parry_b200
    G, ns = _mixed_table(ctx)
    n = 1 << 21
    g, d, s1, p1, s2, p2 = _mixed_pairs(n, ns, seed=32, spread=3.5)
    v1 = (d * (0.5 + g.random((n, 1)) * 3.0) + g.standard_normal((n, 3)) * 0.5).astype(np.float32)
    v2 = (g.standard_normal((n, 3)) * 0.3).astype(np.float32)
    T = lambda x: torch.from_numpy(x).cuda()
    ds1, dp1, ds2, dp2, dv1, dv2 = T(s1), T(p1), T(s2), T(p2), T(v1), T(v2)
    res, out = {}, {}

    def run_cast():
        res["cast"] = parry_b200.cast_shapes(G, ds1, dp1, dv1, ds2, dp2, dv2)
    ms = timed(run_cast, steps=5, warmup=2, flush=flush)
    st = res["cast"][1]
    out["cast_shapes"] = {"value": n / (ms * 1e-3), "unit": "pairs/s", "ms": ms, "hits": float((st == 1).float().mean().item()),
                          "penetrating_starts": float((st == 2).float().mean().item())}

    def run_dist():
        res["dist"] = parry_b200.distance(G, ds1, dp1, ds2, dp2)
    ms = timed(run_dist, steps=5, warmup=2, flush=flush)
    out["distance"] = {"value": n / (ms * 1e-3), "unit": "pairs/s", "ms": ms}

    def run_cp():
        res["cp"] = parry_b200.closest_points(G, ds1, dp1, ds2, dp2, 1.0)
    ms = timed(run_cp, steps=5, warmup=2, flush=flush)
    out["closest_points"] = {"value": n / (ms * 1e-3), "unit": "pairs/s", "ms": ms,
                             "within_margin": float((res["cp"][1] == 1).float().mean().item())}
    # compounds: 4096 of 1-5 parts, 2^20 (compound, shape) pairs
    nc = 4096
    compounds = []
    for c in range(nc):
        k = int(g.integers(1, 6))
        from harness import scenes
        pp = np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.6], axis=1).astype(np.float32)
        compounds.append([(pp[i], int(g.integers(0, ns))) for i in range(k)])
    Cc = parry_b200.Compounds(ctx, G, compounds)
    m = 1 << 20
    cid = T(g.integers(0, nc, m).astype(np.int32))

    def run_comp():
        res["comp"] = Cc.contact_shapes(cid, dp1[:m], ds2[:m], dp2[:m], 0.05)
    ms = timed(run_comp, steps=5, warmup=2, flush=flush)
    out["compound_contacts"] = {"value": m / (ms * 1e-3), "unit": "pairs/s", "ms": ms,
                                "contacts_fraction": float((res["comp"][1] == 1).float().mean().item())}
    out["pairs"] = n
    out["l2"] = "flushed between iterations"
    return out


def also_persistence_and_compounds(ctx, stream, timed, flush, hbm_peak):
    """Two SURVEY §8 (f) paths: manifold persistence (second frame of 2^22 ball / cuboid pairs, half of them drifting by
    2e-4 — most of those keep their manifold — and half by 0.05) and Compound vs Compound contacts (2^20 pairs of 1-5 parts)."""
    import torch
    import parry_b200
    from harness import scenes
    out = {}
    T = lambda x: torch.from_numpy(x).cuda()
    n = 1 << 22
    g = scenes.rng(41)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(0.4), parry_b200.Ball(0.25), parry_b200.Cuboid([0.3, 0.5, 0.4]), parry_b200.Cuboid([0.6, 0.2, 0.2]),
                                parry_b200.Cuboid([0.5, 0.5, 0.5])])
    s1, s2 = g.integers(0, 5, n).astype(np.int32), g.integers(0, 5, n).astype(np.int32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 20], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.3 + 0.2)], axis=1).astype(np.float32)
    p2b = p2.copy()
    p2b[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.5, 2.0e-4, 0.05)).astype(np.float32)
    ds1, ds2, dp1, dp2, dp2b = T(s1), T(s2), T(p1), T(p2), T(p2b)
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, ds1, dp1, ds2, dp2, 0.05, max_points=8)
    ctx.synchronize()
    res = {}

    def run_update():
        res["u"] = parry_b200.contact_manifolds_update(G, ds1, dp1, ds2, dp2b, 0.05, nr, cnt, pts)
    ms = timed(run_update, steps=5, warmup=2, flush=flush)
    kept, cnt2 = res["u"][4], res["u"][1]
    out["manifold_persistence_4M_ball_cuboid_pairs"] = {
        "value": n / (ms * 1e-3), "unit": "pairs/s (second frame: try_update_contacts, recompute the rest, match_contacts)", "ms": ms,
        "kept_fraction_of_nonempty": float((kept[cnt > 0] != 0).float().mean().item()),
        "manifolds_with_points": float((cnt2 > 0).float().mean().item()),
        "note": "includes the wrapper's private copy of last frame's manifolds (1.2 GB)"}
    del res["u"], nr, cnt, pts, st
    G2, ns = _mixed_table(ctx)
    nc = 4096
    compounds = []
    for c in range(nc):
        k = int(g.integers(1, 6))
        pp = np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.6], axis=1).astype(np.float32)
        compounds.append([(pp[i], int(g.integers(0, ns))) for i in range(k)])
    Cc = parry_b200.Compounds(ctx, G2, compounds)
    m = 1 << 20
    c1, c2 = T(g.integers(0, nc, m).astype(np.int32)), T(g.integers(0, nc, m).astype(np.int32))
    q2 = p2[:m].copy()
    q2[:, 4:] = p1[:m, 4:] + d[:m] * (g.random((m, 1)) * 3.0 + 0.2).astype(np.float32)
    dq2 = T(q2)

    def run_cc():
        res["cc"] = Cc.contact_compounds(c1, dp1[:m], c2, dq2, 0.05)
    ms = timed(run_cc, steps=5, warmup=2, flush=flush)
    stc = res["cc"][1]
    out["compound_vs_compound_1M_pairs"] = {"value": m / (ms * 1e-3), "unit": "pairs/s", "ms": ms,
                                            "contacts_fraction": float((stc == 1).float().mean().item()),
                                            "host_fallback": int((stc == 3).sum().item())}
    out["l2"] = "flushed between iterations"
    return out


def also_mixed_sharded(ctx, stream, timed, flush, hbm_peak, comm, dist, rank, world):
    """BASELINE config[4] split over the GPUs of the box (strong scaling: the same 2^21 colliders at every N): every rank computes
    all leaf AABBs and rebuilds the (replicated, deterministic) Bvh, walks the tree only for its interleaved share of the leaves
    (pb2_bvh_self_pairs_shard), runs query::contact on its own pairs, and the compacted contacts (52-byte records + the pair they
    belong to) are gathered on every rank with pb2_comm_allgatherv — all inside the timed region. Cross-rank check after the run: every
    rank holds the same gathered set (checksums), and its size is the same at every N."""
    import torch
    import parry_b200
    from harness import scenes
    n, H = 1 << 21, 4096
    kinds, params, poses, hull_ids = scenes.colliders(n, seed=8, hull_fraction=1.0 / 3.0, n_hulls=H)
    pts, _ = scenes.hull_pool(H, 32, seed=9)
    pts = np.asarray(pts, dtype=np.float32) * 0.5
    tk = np.concatenate([np.full(H, 2, np.uint8), np.where(kinds == 2, 0, kinds).astype(np.uint8)])
    tp = np.concatenate([np.zeros((H, 3), np.float32), params])
    first = np.concatenate([np.arange(H, dtype=np.uint32) * 32, np.zeros(n, np.uint32)])
    count = np.concatenate([np.full(H, 32, np.uint32), np.zeros(n, np.uint32)])
    G = parry_b200.Shapes.from_arrays(ctx, tk, tp, pts.reshape(-1, 3), first, count)
    with torch.cuda.stream(stream):
        cs = torch.from_numpy(np.where(kinds == 2, hull_ids, H + np.arange(n)).astype(np.uint32).view(np.int32)).cuda()
        dposes = torch.from_numpy(poses).cuda()
        all_ids = torch.arange(n, dtype=torch.int32, device="cuda")
    ctx.synchronize()
    aabbs = G.compute_aabbs(cs, dposes)
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    state = {}

    def frame():
        a = G.compute_aabbs(cs, dposes)
        bvh.insert_or_update_partially(a, all_ids, 0.0)
        bvh.rebuild()
        pr = bvh.traverse_bvtt_single_tree_shard(rank, world, capacity=(16 * n) // world + 4096, like=a)
        out, idx = parry_b200.contact_pairs_compact(G, cs, dposes, pr, 0.01)
        with torch.cuda.stream(stream):
            owners = pr[idx.to(torch.int64)]                  # the pair each compacted contact belongs to
        g_out, c1 = comm.allgatherv(out, capacity=state.get("cap", 1 << 22))
        g_pair, c2 = comm.allgatherv(owners, capacity=state.get("cap", 1 << 22))
        state.update(pairs=int(pr.shape[0]), contacts=int(out.shape[0]), g_out=g_out, g_pair=g_pair, counts=c1, cap=int(g_out.shape[0]) + 4096)

    ms = timed(frame, steps=5, warmup=2, flush=flush)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.synchronize()
    g_pair = state["g_pair"].to(torch.int64)
    key = (g_pair[:, 0] << 32) | (g_pair[:, 1] & 0xFFFFFFFF)
    chk = torch.stack([key.sum(), (key * 31 + 7).remainder(1000003).sum(), torch.tensor(int(key.shape[0]), device=key.device)])
    dsum = state["g_out"][:, 12].double().sum().reshape(1)
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).all())
    pairs_total = state["pairs"]
    if world > 1:
        t = torch.tensor([pairs_total], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        pairs_total = int(t.item())
    assert same, "ranks hold different gathered contact sets"
    C_total = int(key.shape[0])
    return {"value": n / (ms * 1e-3), "unit": "colliders/s (frame: AABBs + Bvh rebuild [replicated] + sharded self pairs + contacts + all-gather-v of the contacts)",
            "ms": ms, "n_gpus": world, "scaling": "strong", "colliders": n, "pairs_per_frame": pairs_total, "contacts_per_frame": C_total,
            "contacts_per_rank": state["counts"], "pairs_per_s": pairs_total / (ms * 1e-3),
            "gathered_bytes_per_rank": C_total * 60, "checksum": [int(x) for x in chk.tolist()], "dist_sum": float(dsum.item()),
            "all_ranks_agree": same, "l2": "flushed between iterations",
            "exchange": "pb2_comm_allgatherv (NCCL grouped broadcasts, exact sizes) x2: 52-byte contacts and their 8-byte pair ids"}


MULTI_GPU_ALSO = [("mixed_2M_colliders_sharded", also_mixed_sharded)]   # (name, fn(ctx, stream, timed, flush, hbm_peak, comm, dist, rank, world))

EXTRA_ALSO = [("manifolds_4M_ball_cuboid_pairs", also_manifolds),
              ("sibling_queries_2M_mixed_pairs", also_siblings), ("broadphase_1M_colliders", also_broadphase),
              ("mixed_2M_colliders_pipeline", also_mixed), ("trimesh_contacts_1M_colliders", also_mesh_contacts)]

if __name__ == "__main__":
    try:
        main()
    except BaseException as exc:   # a rank that dies must take the job down at once (peers would wait in a collective for ever)
        if not isinstance(exc, SystemExit) or exc.code not in (0, None):
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
        raise
