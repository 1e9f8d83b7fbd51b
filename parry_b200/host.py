"""Host-side mirror of the reference's API for the hot path, over the C ABI. Names follow parry3d:
`Bvh::from_leaves / refit / intersect_aabb / traverse_bvtt_single_tree / leaf_pairs` (src/partitioning/bvh),
`TriMesh::cast_ray / cast_ray_and_get_normal / cast_local_ray*` (src/query/ray/ray.rs:359-411),
`query::contact` (src/query/contact/contact_shape_shape.rs:123). Every call is the batched form; arrays may be
numpy (host; results come back as numpy) or torch CUDA tensors (device resident; results are CUDA tensors and the
call is asynchronous on the context's stream)."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import MEM_DEVICE, MEM_HOST, Pb2Error, Unsupported

try:  # torch is plumbing only (device memory + streams); the library itself never sees torch types
    import torch
except Exception:  # pragma: no cover
    torch = None


class BvhBuildStrategy:
    """partitioning/bvh/bvh_tree.rs:58-78"""
    Binned = 0
    Ploc = 1


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _prep(x, dtype, mem=None):
    """Returns (keepalive, address, mem) for an input array."""
    if x is None:
        return None, None, mem
    if _is_torch(x):
        tdt = {np.float32: torch.float32, np.uint32: torch.int32, np.uint8: torch.uint8}[dtype]
        x0 = x
        if x.dtype != tdt:
            if dtype is np.uint32 and x.dtype in (torch.int64, torch.uint32):
                x = x.to(torch.int32)
            else:
                x = x.to(tdt)
        x = x.contiguous()
        if x.is_cuda:
            if x is not x0:
                # a conversion kernel ran on torch's current stream; the library's stream is non-blocking and would not wait for it
                torch.cuda.current_stream(x.device).synchronize()
            return x, x.data_ptr(), MEM_DEVICE
        x = x.numpy()
    a = np.ascontiguousarray(x, dtype=dtype)
    return a, a.ctypes.data, MEM_HOST


def _empty(shape, dtype, mem, device):
    if mem == MEM_DEVICE:
        tdt = {np.float32: torch.float32, np.uint32: torch.int32, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
        t = torch.empty(shape, dtype=tdt, device=device)
        return t, t.data_ptr()
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data


class Context:
    """One CUDA device + one stream (pb2_ctx). With `stream=` (a torch.cuda.Stream) work is enqueued there.
    The context's own stream is non-blocking: device tensors handed to a call must already be complete with respect to it
    (produce them under `torch.cuda.stream(ctx.torch_stream())`, or synchronize the producing stream first), and results are
    ordered on it (`ctx.synchronize()` before reading them from another stream)."""

    def __init__(self, device=0, stream=None):
        self._lib = _ffi.lib()
        self.device = int(device)
        h = C.c_void_p()
        if stream is not None:
            st = self._lib.pb2_ctx_create_on_stream(self.device, C.c_void_p(stream.cuda_stream), C.byref(h))
        else:
            st = self._lib.pb2_ctx_create(self.device, C.byref(h))
        if st != 0:
            raise Pb2Error(st, "pb2_ctx_create failed: no usable CUDA device %d (no CPU fallback)" % self.device)
        self.h = h
        self.torch_device = "cuda:%d" % self.device

    def check(self, st):
        if st == 0:
            return
        msg = self._lib.pb2_last_error(self.h)
        msg = msg.decode() if msg else ""
        if st == _ffi.PB2_ERR_UNSUPPORTED:
            raise Unsupported(st, msg)
        raise Pb2Error(st, msg)

    def synchronize(self):
        self.check(self._lib.pb2_ctx_synchronize(self.h))

    def enable_phase_timing(self, on=True):
        self.check(self._lib.pb2_ctx_enable_phase_timing(self.h, 1 if on else 0))

    def contact_phase_times(self):
        """(gjk_ms, epa_ms, finish_ms, epa_runs) of the last contact-family call (CUDA events on the library's stream)."""
        a, b, c, r = C.c_float(0), C.c_float(0), C.c_float(0), C.c_uint64(0)
        self.check(self._lib.pb2_contact_phase_times(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(r)))
        return a.value, b.value, c.value, int(r.value)

    @property
    def stream_ptr(self):
        return self._lib.pb2_ctx_stream(self.h)

    def torch_stream(self):
        return torch.cuda.ExternalStream(self.stream_ptr, device=self.torch_device)

    @property
    def launch_count(self):
        return int(self._lib.pb2_ctx_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.pb2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Bvh:
    """partitioning::Bvh"""

    def __init__(self, ctx, handle, owned=True, keep=None):
        self.ctx, self.h, self.owned, self._keep = ctx, handle, owned, keep

    @staticmethod
    def from_leaves(ctx, strategy, aabbs):
        """Bvh::from_leaves (bvh_tree.rs:1835). aabbs: (n, 6) [mins, maxs]."""
        n = 0 if aabbs is None else int(aabbs.shape[0])
        keep, ptr, mem = _prep(aabbs, np.float32)
        h = C.c_void_p()
        ctx.check(ctx._lib.pb2_bvh_build(ctx.h, ptr, n, int(strategy), mem if mem is not None else MEM_HOST, C.byref(h)))
        return Bvh(ctx, h)

    def leaf_count(self):
        return int(self.ctx._lib.pb2_bvh_leaf_count(self.h))

    def node_count(self):
        return int(self.ctx._lib.pb2_bvh_node_count(self.h))

    def is_empty(self):
        return self.leaf_count() == 0

    def root_aabb(self):
        out = np.empty(6, dtype=np.float32)
        self.ctx.check(self.ctx._lib.pb2_bvh_root_aabb(self.ctx.h, self.h, out.ctypes.data))
        return out

    def insert_or_update_partially(self, aabbs, leaf_indices=None, change_detection_margin=0.0):
        """Batched Bvh::insert_or_update_partially for existing leaves (bvh_insert.rs:209-231)."""
        k1, p1, mem = _prep(aabbs, np.float32)
        k2, p2, _ = _prep(leaf_indices, np.uint32, mem)
        self.ctx.check(self.ctx._lib.pb2_bvh_update_leaves(self.ctx.h, self.h, p2, p1, int(aabbs.shape[0]),
                                                          float(change_detection_margin), mem))

    def insert(self, aabbs, leaf_indices):
        """Batched Bvh::insert (bvh_insert.rs:126-197): new leaf indices grow the tree, known ones are updated; the
        tree is rebuilt (structural edits are whole-tree rebuilds on the GPU)."""
        ids = leaf_indices
        top = int(ids.max().item() if _is_torch(ids) else np.asarray(ids).max()) + 1 if int(aabbs.shape[0]) else 0
        if top > self.leaf_count():
            self.ctx.check(self.ctx._lib.pb2_bvh_resize(self.ctx.h, self.h, top))
        self.insert_or_update_partially(aabbs, leaf_indices, 0.0)
        self.rebuild()

    def remove(self, leaf_indices):
        """Batched Bvh::remove (bvh_tree.rs:2360-2427)."""
        k, p, mem = _prep(leaf_indices, np.uint32)
        self.ctx.check(self.ctx._lib.pb2_bvh_remove_leaves(self.ctx.h, self.h, p, int(leaf_indices.shape[0]), mem))

    @staticmethod
    def from_iter(ctx, strategy, leaves):
        """Bvh::from_iter (bvh_tree.rs:1891-1955): `leaves` yields (index, aabb); indices need not be dense — the gaps become
        removed leaves (inert slots), as after Bvh::remove."""
        items = [(int(i), np.asarray(a, dtype=np.float32).reshape(6)) for i, a in leaves]
        n = max((i for i, _ in items), default=-1) + 1
        aabbs = np.zeros((n, 6), dtype=np.float32)
        present = np.zeros(n, dtype=bool)
        for i, a in items:
            aabbs[i] = a
            present[i] = True
        bvh = Bvh.from_leaves(ctx, strategy, aabbs)
        if n and not present.all():
            bvh.remove(np.nonzero(~present)[0].astype(np.uint32))
            bvh.rebuild(strategy)
        return bvh

    def refit(self):
        """Bvh::refit (bvh_refit.rs:170-320): bottom-up AABB update after insert_or_update_partially."""
        self.ctx.check(self.ctx._lib.pb2_bvh_refit(self.ctx.h, self.h))

    def refit_without_opt(self):
        """Bvh::refit_without_opt (bvh_refit.rs:326-375): the same pass here (the node layout is rewritten by rebuild only)."""
        self.refit()

    def optimize_incremental(self):
        """Bvh::optimize_incremental (bvh_optimize.rs:237-326) re-bins ~5 % of the leaves per frame to keep a refitted tree
        from degrading; a full LBVH rebuild costs ~1 ms per million leaves on this path, so the whole tree is rebuilt."""
        self.rebuild()

    def rebuild(self, strategy=BvhBuildStrategy.Binned):
        self.ctx.check(self.ctx._lib.pb2_bvh_rebuild(self.ctx.h, self.h, int(strategy)))

    def download(self):
        """(nodes, parents, leaf_node_indices): nodes as a structured array in the BvhNodeWide layout."""
        nn, nl = self.node_count(), self.leaf_count()
        nodes = np.zeros(nn, dtype=NODE_WIDE_DTYPE)
        parents = np.zeros(nn, dtype=np.uint32)
        leaf_idx = np.zeros(nl, dtype=np.uint32)
        if nn:
            self.ctx.check(self.ctx._lib.pb2_bvh_download(self.ctx.h, self.h, nodes.ctypes.data, parents.ctypes.data,
                                                         leaf_idx.ctypes.data, MEM_HOST))
        return nodes, parents, leaf_idx

    def intersect_aabb(self, queries, capacity=None):
        """Batched Bvh::intersect_aabb (bvh_queries.rs:203). Returns (offsets[m+1], leaf_ids[count])."""
        m = int(queries.shape[0])
        kq, pq, mem = _prep(queries, np.float32)
        dev = self.ctx.torch_device
        cap = int(capacity) if capacity is not None else max(1024, 16 * m)
        while True:
            offs, po = _empty((m + 1,), np.uint32, mem, dev)
            ids, pi = _empty((cap,), np.uint32, mem, dev)
            cnt = C.c_uint64(0)
            st = self.ctx._lib.pb2_bvh_intersect_aabbs(self.ctx.h, self.h, pq, m, po, pi, cap, C.byref(cnt), mem)
            if st == _ffi.PB2_ERR_OVERFLOW:
                cap = int(cnt.value)
                continue
            self.ctx.check(st)
            return offs, ids[: int(cnt.value)]

    def traverse_bvtt_single_tree(self, change_detection=False, capacity=None, like=None):
        """Bvh::traverse_bvtt_single_tree (bvh_traverse_bvtt.rs:19): all overlapping leaf pairs, (count, 2)."""
        mem = MEM_DEVICE if (like is not None and _is_torch(like) and like.is_cuda) else MEM_HOST
        cap = int(capacity) if capacity is not None else max(1024, 8 * self.leaf_count())
        while True:
            pairs, pp = _empty((cap, 2), np.uint32, mem, self.ctx.torch_device)
            cnt = C.c_uint64(0)
            st = self.ctx._lib.pb2_bvh_self_pairs(self.ctx.h, self.h, int(bool(change_detection)), pp, cap, C.byref(cnt), mem)
            if st == _ffi.PB2_ERR_OVERFLOW:
                cap = int(cnt.value)
                continue
            self.ctx.check(st)
            return pairs[: int(cnt.value)]

    def traverse_bvtt_single_tree_shard(self, shard, n_shards, change_detection=False, capacity=None, like=None):
        """This rank's part of traverse_bvtt_single_tree when the broad phase is split over n_shards GPUs with the Bvh replicated
        (SURVEY 8e): the leaves at sorted positions p = shard (mod n_shards) walk the tree; the shards' outputs partition the pair set."""
        mem = MEM_DEVICE if (like is not None and _is_torch(like) and like.is_cuda) else MEM_HOST
        cap = int(capacity) if capacity is not None else max(1024, 8 * self.leaf_count() // int(n_shards) + 1024)
        while True:
            pairs, pp = _empty((cap, 2), np.uint32, mem, self.ctx.torch_device)
            cnt = C.c_uint64(0)
            st = self.ctx._lib.pb2_bvh_self_pairs_shard(self.ctx.h, self.h, int(bool(change_detection)), int(shard), int(n_shards), pp, cap,
                                                        C.byref(cnt), mem)
            if st == _ffi.PB2_ERR_OVERFLOW:
                cap = int(cnt.value)
                continue
            self.ctx.check(st)
            return pairs[: int(cnt.value)]

    def leaf_pairs(self, other, capacity=None, like=None):
        """Bvh::leaf_pairs(other, |a, b| a.intersects(b)) (bvh_traverse_bvtt.rs:210)."""
        mem = MEM_DEVICE if (like is not None and _is_torch(like) and like.is_cuda) else MEM_HOST
        cap = int(capacity) if capacity is not None else max(1024, 8 * max(self.leaf_count(), other.leaf_count()))
        while True:
            pairs, pp = _empty((cap, 2), np.uint32, mem, self.ctx.torch_device)
            cnt = C.c_uint64(0)
            st = self.ctx._lib.pb2_bvh_leaf_pairs(self.ctx.h, self.h, other.h, pp, cap, C.byref(cnt), mem)
            if st == _ffi.PB2_ERR_OVERFLOW:
                cap = int(cnt.value)
                continue
            self.ctx.check(st)
            return pairs[: int(cnt.value)]

    def cast_ray(self, shapes, shape_ids, poses, rays, max_time_of_impact, solid=True, with_normal=False):
        """Batched Bvh::cast_ray with typed Ball/Cuboid leaves (bvh_queries.rs:260)."""
        m = int(rays.shape[0])
        kr, pr, mem = _prep(rays, np.float32)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        kp, pp, _ = _prep(poses, np.float32, mem)
        dev = self.ctx.torch_device
        toi, pt = _empty((m,), np.float32, mem, dev)
        leaf, pl = _empty((m,), np.uint32, mem, dev)
        normal = feature = None
        pn = pf = None
        if with_normal:
            normal, pn = _empty((m, 3), np.float32, mem, dev)
            feature, pf = _empty((m,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_bvh_cast_rays_shapes(self.ctx.h, self.h, shapes.h, ps, pp, pr, m, float(max_time_of_impact),
                                                             int(bool(solid)), pt, pl, pn, pf, mem))
        return (toi, leaf, normal, feature) if with_normal else (toi, leaf)

    def project_point(self, shapes, shape_ids, poses, points, max_distance, solid=True):
        """Batched Bvh::project_point with typed leaves (bvh_queries.rs:213-227): leaf i = shapes[shape_ids[i]] at poses[i]. Returns
        (proj (m, 3) world space, inside (m,), leaf (m,), status (m,): 0 nothing within max_distance, 1 found, 3 host)."""
        m = int(points.shape[0])
        kq, pq, mem = _prep(points, np.float32)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        kp, pp, _ = _prep(poses, np.float32, mem)
        dev = self.ctx.torch_device
        proj, ppr = _empty((m, 3), np.float32, mem, dev)
        inside, pin = _empty((m,), np.uint8, mem, dev)
        leaf, pl = _empty((m,), np.uint32, mem, dev)
        status, pst = _empty((m,), np.uint8, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_bvh_project_points_shapes(self.ctx.h, self.h, shapes.h, ps, pp, pq, m, float(max_distance),
                                                                   int(bool(solid)), ppr, pin, pl, pst, mem))
        return proj, inside, leaf, status

    def close(self):
        if self.owned and self.h:
            self.ctx._lib.pb2_bvh_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


NODE_HALF_FIELDS = [("mins", np.float32, (3,)), ("children", np.uint32), ("maxs", np.float32, (3,)), ("data", np.uint32)]
NODE_WIDE_DTYPE = np.dtype([("left", np.dtype(NODE_HALF_FIELDS)), ("right", np.dtype(NODE_HALF_FIELDS))])
assert NODE_WIDE_DTYPE.itemsize == 64


class TriMesh:
    """shape::TriMesh (vertices + indices + Bvh) with the RayCast impl (query/ray/ray_trimesh.rs)."""

    def __init__(self, ctx, vertices, indices):
        """TriMesh::new (shape/trimesh.rs:607)."""
        self.ctx = ctx
        kv, pv, mem = _prep(vertices, np.float32)
        ki, pi, _ = _prep(indices, np.uint32, mem)
        self.num_vertices = int(vertices.shape[0])
        self.num_triangles = int(indices.shape[0])
        h = C.c_void_p()
        ctx.check(ctx._lib.pb2_trimesh_create(ctx.h, pv, self.num_vertices, pi, self.num_triangles, mem, C.byref(h)))
        self.h = h

    def wide_bytes(self):
        """Bytes of the traversal arrays (wide nodes + pre-gathered triangles) a ray cast walks."""
        return int(self.ctx._lib.pb2_trimesh_traversal_bytes(self.h))

    def bvh(self):
        return Bvh(self.ctx, C.c_void_p(self.ctx._lib.pb2_trimesh_bvh(self.h)), owned=False, keep=self)

    def _cast(self, pose, rays, max_toi, solid, with_normal, out=None, culling=0):
        m = int(rays.shape[0])
        kr, pr, mem = _prep(rays, np.float32)
        kp, pp, _ = _prep(pose, np.float32, mem)
        dev = self.ctx.torch_device
        if out is not None:
            toi, tri = out[0], out[1]
            pt = toi.data_ptr() if _is_torch(toi) else toi.ctypes.data
            pl = tri.data_ptr() if _is_torch(tri) else tri.ctypes.data
        else:
            toi, pt = _empty((m,), np.float32, mem, dev)
            tri, pl = _empty((m,), np.uint32, mem, dev)
        normal = feature = None
        pn = pf = None
        if with_normal:
            if out is not None and len(out) == 4:
                normal, feature = out[2], out[3]
                pn = normal.data_ptr() if _is_torch(normal) else normal.ctypes.data
                pf = feature.data_ptr() if _is_torch(feature) else feature.ctypes.data
            else:
                normal, pn = _empty((m, 3), np.float32, mem, dev)
                feature, pf = _empty((m,), np.uint32, mem, dev)
        if culling:
            self.ctx.check(self.ctx._lib.pb2_trimesh_cast_rays_with_culling(self.ctx.h, self.h, pp, pr, m, float(max_toi), int(culling),
                                                                           pt, pl, pn, pf, mem))
        else:
            self.ctx.check(self.ctx._lib.pb2_trimesh_cast_rays(self.ctx.h, self.h, pp, pr, m, float(max_toi), int(bool(solid)),
                                                              pt, pl, pn, pf, mem))
        return (toi, tri, normal, feature) if with_normal else (toi, tri)

    IGNORE_BACKFACES, IGNORE_FRONTFACES = 1, 2  # RayCullingMode (ray_trimesh.rs:50-56)

    def cast_ray_with_culling(self, m, rays, max_time_of_impact, culling, out=None):
        """TriMesh::cast_ray_with_culling (ray_trimesh.rs:139-150): (toi, tri, normal, feature)."""
        return self._cast(m, rays, max_time_of_impact, True, True, out, culling)

    def cast_local_ray_with_culling(self, rays, max_time_of_impact, culling, out=None):
        """TriMesh::cast_local_ray_with_culling (ray_trimesh.rs:155-178)."""
        return self._cast(None, rays, max_time_of_impact, True, True, out, culling)

    def cast_ray(self, m, rays, max_time_of_impact, solid=True, out=None):
        """RayCast::cast_ray (ray.rs:381-390), batched: returns (toi, tri); tri == INVALID_U32 means None."""
        return self._cast(m, rays, max_time_of_impact, solid, False, out)

    def cast_ray_and_get_normal(self, m, rays, max_time_of_impact, solid=True, out=None):
        """RayCast::cast_ray_and_get_normal (ray.rs:393-404): (toi, tri, normal, feature)."""
        return self._cast(m, rays, max_time_of_impact, solid, True, out)

    def intersects_ray(self, m, rays, max_time_of_impact):
        """RayCast::intersects_ray (ray.rs:403-411), batched: (n,) bool."""
        toi, tri = self.cast_ray(m, rays, max_time_of_impact, True)
        return (tri != _ffi.INVALID_U32) if not _is_torch(tri) else (tri != -1)

    def cast_local_ray(self, rays, max_time_of_impact, solid=True, out=None):
        return self._cast(None, rays, max_time_of_impact, solid, False, out)

    def cast_local_ray_and_get_normal(self, rays, max_time_of_impact, solid=True, out=None):
        return self._cast(None, rays, max_time_of_impact, solid, True, out)

    def cast_local_ray_allgather(self, rays, max_time_of_impact, peer_toi_ptrs, peer_tri_ptrs, rank, elem_offset, chunks=4):
        """Range-split batch on several GPUs: casts this rank's (device-resident) rays and pushes the results into every
        rank's gather buffers over NVLink while traversing (pb2_trimesh_cast_rays_allgather). peer_*_ptrs: ctypes arrays of
        device pointers, one per rank (parry_b200.sharding.PeerHitGather builds them from symmetric memory)."""
        m = int(rays.shape[0])
        kr, pr, mem = _prep(rays, np.float32)
        if mem != MEM_DEVICE:
            raise ValueError("cast_local_ray_allgather needs device-resident rays")
        self.ctx.check(self.ctx._lib.pb2_trimesh_cast_rays_allgather(self.ctx.h, self.h, None, pr, m, float(max_time_of_impact),
                                                                    peer_toi_ptrs, peer_tri_ptrs, len(peer_toi_ptrs), int(rank),
                                                                    int(elem_offset), int(chunks)))

    def project_point(self, m, points, solid=True):
        """PointQuery::project_point(m, pt, solid), batched: (proj (n, 3), is_inside (n,) u8, triangle (n,) u32)."""
        n = int(points.shape[0])
        kp, pp, mem = _prep(points, np.float32)
        km, pm, _ = _prep(m, np.float32, mem)
        dev = self.ctx.torch_device
        proj, ppr = _empty((n, 3), np.float32, mem, dev)
        inside, pin = _empty((n,), np.uint8, mem, dev)
        tri, ptr = _empty((n,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_trimesh_project_points(self.ctx.h, self.h, pm, pp, n, int(bool(solid)), ppr, pin, ptr, mem))
        return proj, inside, tri

    def project_local_point(self, points, solid=True):
        return self.project_point(None, points, solid)

    def contact_shapes(self, mesh_pose, shapes, shape_ids, poses, prediction):
        """query::contact(mesh_pose, self, poses[k], shapes[shape_ids[k]], prediction) for every k (the composite-shape arm,
        contact_composite_shape_shape.rs:14-61). Returns (contacts (n, 13), status (n,), part (n,) winning triangle)."""
        n = int(poses.shape[0])
        kp, pp, mem = _prep(poses, np.float32)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        km, pm, _ = _prep(mesh_pose, np.float32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        part, ppart = _empty((n,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_trimesh_contact_shapes(self.ctx.h, self.h, pm, shapes.h, ps, pp, n, float(prediction), po, pst, ppart, mem))
        return out, status, part

    def cast_shapes(self, mesh_pose, mesh_vel, shapes, shape_ids, poses, vels, options=None, mesh_second=False):
        """query::cast_shapes(mesh_pose, mesh_vel, self, poses[k], vels[k], shapes[shape_ids[k]], options) for every k — or, with
        mesh_second, the shape as shape 1 and the swapped hit (shape_cast_composite_shape_shape.rs:65-105). Returns (hits (n, 13) as
        parry_b200.cast_shapes, status (n,), part (n,) = the triangle hit)."""
        o = options or ShapeCastOptions()
        n = int(poses.shape[0])
        kp, pp, mem = _prep(poses, np.float32)
        kv, pv, _ = _prep(vels, np.float32, mem)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        km, pm, _ = _prep(mesh_pose, np.float32, mem)
        kmv, pmv, _ = _prep(mesh_vel, np.float32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        part, ppart = _empty((n,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_trimesh_cast_shapes(self.ctx.h, self.h, pm, pmv, shapes.h, ps, pp, pv, int(mesh_second),
                                                             o.max_time_of_impact, o.target_distance, int(o.stop_at_penetration),
                                                             int(o.compute_impact_geometry_on_penetration), n, po, pst, ppart, mem))
        return out, status, part

    def distance_shapes(self, mesh_pose, shapes, shape_ids, poses, mesh_second=False):
        """query::distance(mesh_pose, self, poses[k], shapes[shape_ids[k]]) for every k, or with mesh_second the shape first
        (distance_composite_shape_shape.rs:46-77). Returns (dist (n,), status (n,), part (n,) = the closest triangle)."""
        n = int(poses.shape[0])
        kp, pp, mem = _prep(poses, np.float32)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        km, pm, _ = _prep(mesh_pose, np.float32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n,), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        part, ppart = _empty((n,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_trimesh_distance_shapes(self.ctx.h, self.h, pm, shapes.h, ps, pp, int(mesh_second), n, po, pst, ppart, mem))
        return out, status, part

    def cast_trimesh(self, poses1, vels1, other, poses2, vels2, options=None):
        """query::cast_shapes(poses1[k], vels1[k], self, poses2[k], vels2[k], other, options) for every k (two TriMeshes; the
        reference's trimesh_trimesh_toi.rs). Returns (hits (n, 13), status (n,), parts (n, 2) = the triangle of each mesh)."""
        o = options or ShapeCastOptions()
        n = int(poses1.shape[0])
        k1, p1, mem = _prep(poses1, np.float32)
        k2, p2, _ = _prep(poses2, np.float32, mem)
        kv1, v1, _ = _prep(vels1, np.float32, mem)
        kv2, v2, _ = _prep(vels2, np.float32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        parts, pp = _empty((n, 2), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_trimesh_cast_trimesh(self.ctx.h, self.h, p1, v1, other.h, p2, v2, o.max_time_of_impact, o.target_distance,
                                                              int(o.stop_at_penetration), int(o.compute_impact_geometry_on_penetration), n, po,
                                                              pst, pp, mem))
        return out, status, parts

    def close(self):
        if self.h:
            self.ctx._lib.pb2_trimesh_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ball:
    """shape::Ball (shape/ball.rs:50)"""
    kind = 0

    def __init__(self, radius):
        self.radius = float(radius)


class Cuboid:
    """shape::Cuboid (shape/cuboid.rs:69)"""
    kind = 1

    def __init__(self, half_extents):
        self.half_extents = np.asarray(half_extents, dtype=np.float32)


class ConvexPolyhedron:
    """shape::ConvexPolyhedron — only `points()` matters on this path (shape/convex_polyhedron.rs:172-185)."""
    kind = 2

    def __init__(self, points):
        self.points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)


class Shapes:
    """A table of shapes living on the device; batches reference shapes by index (pb2_shapes)."""

    def __init__(self, ctx, shapes):
        self.ctx = ctx
        n = len(shapes)
        kinds = np.zeros(n, dtype=np.uint8)
        params = np.zeros((n, 4), dtype=np.float32)
        pu = params.view(np.uint32)
        pts = []
        npts = 0
        for i, s in enumerate(shapes):
            kinds[i] = s.kind
            if s.kind == 0:
                params[i, 0] = s.radius
            elif s.kind == 1:
                params[i, :3] = s.half_extents
            else:
                pu[i, 0] = npts
                pu[i, 1] = len(s.points)
                pts.append(s.points)
                npts += len(s.points)
        points = np.concatenate(pts, axis=0).astype(np.float32) if pts else np.zeros((0, 3), dtype=np.float32)
        self.kinds, self.params, self.points = kinds, params, np.ascontiguousarray(points)
        self.n = n
        h = C.c_void_p()
        ctx.check(ctx._lib.pb2_shapes_create(ctx.h, kinds.ctypes.data, params.ctypes.data, n,
                                             self.points.ctypes.data if npts else None, npts, C.byref(h)))
        self.h = h

    @classmethod
    def from_arrays(cls, ctx, kinds, params, hull_points=None, hull_first=None, hull_count=None):
        """Bulk constructor: kinds (n,) u8 (0 ball, 1 cuboid, 2 convex), params (n, 3): radius / half extents; convex
        shape i uses hull_points[hull_first[i] : hull_first[i] + hull_count[i]]."""
        self = cls.__new__(cls)
        self.ctx = ctx
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        n = len(kinds)
        p4 = np.zeros((n, 4), dtype=np.float32)
        p4[:, :3] = np.asarray(params, dtype=np.float32).reshape(n, -1)[:, :3]
        cv = kinds == 2
        if cv.any():
            pu = p4.view(np.uint32)
            pu[cv, 0] = np.asarray(hull_first, dtype=np.uint32)[cv]
            pu[cv, 1] = np.asarray(hull_count, dtype=np.uint32)[cv]
            p4[cv, 2] = 0.0
        points = np.ascontiguousarray(hull_points, dtype=np.float32).reshape(-1, 3) if hull_points is not None else np.zeros((0, 3), np.float32)
        self.kinds, self.params, self.points, self.n = kinds, p4, points, n
        h = C.c_void_p()
        ctx.check(ctx._lib.pb2_shapes_create(ctx.h, kinds.ctypes.data, p4.ctypes.data, n,
                                             points.ctypes.data if len(points) else None, len(points), C.byref(h)))
        self.h = h
        return self

    def set_hull_topology(self, topology):
        """Face topology of the ConvexPolyhedron entries (what ConvexPolyhedron::faces() / vertices_adj_to_face() /
        edges_adj_to_face() hold in parry): dict(hull_face_first, hull_face_count (per table entry), face_normal (nf, 3),
        face_first, face_count (nf), vertices_adj_to_face, edges_adj_to_face). Enables the pfm_pfm arm of contact_manifolds."""
        t = {k: np.ascontiguousarray(topology[k], dtype=np.float32 if k == "face_normal" else np.uint32) for k in
             ("hull_face_first", "hull_face_count", "face_normal", "face_first", "face_count", "vertices_adj_to_face", "edges_adj_to_face")}
        assert len(t["hull_face_first"]) == self.n and len(t["hull_face_count"]) == self.n
        self.ctx.check(self.ctx._lib.pb2_shapes_set_hull_topology(
            self.ctx.h, self.h, t["hull_face_first"].ctypes.data, t["hull_face_count"].ctypes.data, t["face_normal"].ctypes.data,
            t["face_first"].ctypes.data, t["face_count"].ctypes.data, len(t["face_first"]), t["vertices_adj_to_face"].ctypes.data,
            t["edges_adj_to_face"].ctypes.data, len(t["vertices_adj_to_face"])))
        if "vert_first" in topology:   # vertex side (ball-vs-hull arm): per-point adjacency, edge directions
            v = {k: np.ascontiguousarray(topology[k], dtype=np.float32 if k == "edge_dir" else np.uint32) for k in
                 ("vert_first", "vert_count", "faces_adj_to_vertex", "edges_adj_to_vertex", "hull_edge_first", "edge_dir")}
            assert len(v["hull_edge_first"]) == self.n
            self.ctx.check(self.ctx._lib.pb2_shapes_set_hull_vertex_topology(
                self.ctx.h, self.h, v["vert_first"].ctypes.data, v["vert_count"].ctypes.data, v["faces_adj_to_vertex"].ctypes.data,
                v["edges_adj_to_vertex"].ctypes.data, len(v["faces_adj_to_vertex"]), v["hull_edge_first"].ctypes.data, v["edge_dir"].ctypes.data,
                len(v["edge_dir"])))

    def compute_aabbs(self, shape_ids, poses):
        """Shape::compute_aabb(pos), batched (shape/shape.rs:369)."""
        n = int(poses.shape[0])
        kp, pp, mem = _prep(poses, np.float32)
        ks, ps, _ = _prep(shape_ids, np.uint32, mem)
        out, po = _empty((n, 6), np.float32, mem, self.ctx.torch_device)
        self.ctx.check(self.ctx._lib.pb2_shapes_compute_aabbs(self.ctx.h, self.h, ps, pp, n, po, mem))
        return out

    def close(self):
        if self.h:
            self.ctx._lib.pb2_shapes_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Compounds:
    """A table of Compound shapes (shape/compound.rs:113-144) over a Shapes table: compound c = the parts
    [(pose, shapes[shape_id]), ...] given as `compounds[c]`. `contact_shapes` is query::contact with the compound on one side."""

    def __init__(self, ctx, shapes, compounds):
        self.ctx, self.shapes = ctx, shapes
        first, count, part_shape, part_pose = [], [], [], []
        for parts in compounds:
            first.append(len(part_shape))
            count.append(len(parts))
            for pose, sid in parts:
                part_pose.append(np.asarray(pose, dtype=np.float32).reshape(7))
                part_shape.append(int(sid))
        self.first = np.asarray(first, dtype=np.uint32)
        self.count = np.asarray(count, dtype=np.uint32)
        self.part_shape = np.asarray(part_shape, dtype=np.uint32)
        self.part_pose = np.ascontiguousarray(np.stack(part_pose) if part_pose else np.zeros((0, 7)), dtype=np.float32)
        h = C.c_void_p()
        ctx.check(ctx._lib.pb2_compounds_create(ctx.h, shapes.h, self.first.ctypes.data, self.count.ctypes.data, len(self.first),
                                                self.part_shape.ctypes.data, self.part_pose.ctypes.data, len(self.part_shape), C.byref(h)))
        self.h = h

    def contact_shapes(self, compound_ids, compound_poses, shape_ids, shape_poses, prediction, compound_second=False):
        """query::contact(compound_poses[k], compound k, shape_poses[k], shape k, prediction) — or, with compound_second,
        query::contact(shape_poses[k], shape k, compound_poses[k], compound k, prediction) — for every k
        (contact_composite_shape_shape.rs:14-76). Returns (contacts (n, 13), status (n,), part (n,) winning part index)."""
        n = int(compound_poses.shape[0])
        kc, pc, mem = _prep(compound_poses, np.float32)
        ks, ps, _ = _prep(shape_poses, np.float32, mem)
        kci, pci, _ = _prep(compound_ids, np.uint32, mem)
        ksi, psi, _ = _prep(shape_ids, np.uint32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        part, ppart = _empty((n,), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_compound_contact_shapes(self.ctx.h, self.h, pci, pc, psi, ps, n, float(prediction), int(compound_second),
                                                                 po, pst, ppart, mem))
        return out, status, part

    def contact_compounds(self, ids1, poses1, ids2, poses2, prediction):
        """query::contact(poses1[k], compound ids1[k], poses2[k], compound ids2[k], prediction) for every k (the composite arm of
        DefaultQueryDispatcher::contact nested through contact_shape_composite_shape). Returns (contacts (n, 13), status (n,),
        parts (n, 2) winning part of each compound)."""
        n = int(poses1.shape[0])
        k1, p1, mem = _prep(poses1, np.float32)
        k2, p2, _ = _prep(poses2, np.float32, mem)
        ki1, pi1, _ = _prep(ids1, np.uint32, mem)
        ki2, pi2, _ = _prep(ids2, np.uint32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        parts, pp = _empty((n, 2), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_compound_contact_compounds(self.ctx.h, self.h, pi1, p1, pi2, p2, n, float(prediction), po, pst, pp, mem))
        return out, status, parts

    def contact_trimesh(self, ids, poses, mesh, mesh_pose, prediction, mesh_first=False):
        """query::contact(poses[k], compound ids[k], mesh_pose, mesh, prediction) for every k (the composite arm with the Compound
        first, contact_composite_shape_shape.rs:12-48 over the parts and :63-76 over the triangles) — or, with mesh_first,
        query::contact(mesh_pose, mesh, poses[k], compound ids[k], prediction). Returns (contacts (n, 13), status (n,), parts (n, 2) =
        {winning part, winning triangle})."""
        n = int(poses.shape[0])
        kp, pp, mem = _prep(poses, np.float32)
        ki, pi, _ = _prep(ids, np.uint32, mem)
        km, pm, _ = _prep(mesh_pose, np.float32, mem)
        dev = self.ctx.torch_device
        out, po = _empty((n, 13), np.float32, mem, dev)
        status, pst = _empty((n,), np.uint8, mem, dev)
        parts, ppart = _empty((n, 2), np.uint32, mem, dev)
        self.ctx.check(self.ctx._lib.pb2_compound_contact_trimesh(self.ctx.h, self.h, pi, pp, mesh.h, pm, n, float(prediction), int(mesh_first),
                                                                  po, pst, ppart, mem))
        return out, status, parts

    def close(self):
        if self.h:
            self.ctx._lib.pb2_compounds_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """pb2_comm: the multi-GPU exchange of the C ABI (NCCL underneath), one per Context / process. `id128` comes from
    Comm.unique_id() on rank 0 and reaches the other ranks through the host's own transport; from_torch_distributed() uses an
    initialised torch.distributed group (any backend) for exactly that and nothing else."""

    def __init__(self, ctx, id128, rank, nranks):
        self.ctx, self.rank, self.nranks = ctx, int(rank), int(nranks)
        h = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128))
        ctx.check(ctx._lib.pb2_comm_create(ctx.h, buf, self.rank, self.nranks, C.byref(h)))
        self.h = h

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        st = _ffi.lib().pb2_comm_unique_id(buf)
        if st != 0:
            raise Pb2Error("pb2_comm_unique_id failed (%d): is libnccl.so.2 present?" % st)
        return bytes(buf.raw)

    @staticmethod
    def from_torch_distributed(ctx, group=None):
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return Comm(ctx, box[0], rank, world)

    def allgather(self, send, recv):
        """Fixed-size all-gather of device tensors: recv = nranks x send (rank-major)."""
        nbytes = send.numel() * send.element_size()
        assert recv.numel() * recv.element_size() == nbytes * self.nranks
        self.ctx.check(self.ctx._lib.pb2_comm_allgather(self.h, send.data_ptr(), recv.data_ptr(), nbytes))
        return recv

    def allgather_counts(self, mine):
        out = (C.c_uint64 * self.nranks)()
        self.ctx.check(self.ctx._lib.pb2_comm_allgather_counts(self.h, int(mine), out))
        return [int(x) for x in out]

    def allgatherv(self, rows, capacity=None):
        """Variable-size gather of compacted records (rows: (count, ...) device tensor): (all rows in rank order, counts per rank)."""
        count = int(rows.shape[0])
        rows = rows.contiguous()
        elem = rows.element_size() * (int(np.prod(rows.shape[1:])) if rows.dim() > 1 else 1)
        cap = int(capacity) if capacity is not None else max(1, count * self.nranks * 2)
        counts = (C.c_uint64 * self.nranks)()
        total = C.c_uint64(0)
        while True:
            out = torch.empty((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
            st = self.ctx._lib.pb2_comm_allgatherv(self.h, rows.data_ptr() if count else None, count, elem, out.data_ptr(), cap, counts, C.byref(total))
            if st == _ffi.PB2_ERR_OVERFLOW:
                # every rank saw the same total and takes the same branch: the retry stays collective
                cap = int(total.value)
                continue
            self.ctx.check(st)
            return out[: int(total.value)], [int(x) for x in counts]

    def barrier(self):
        self.ctx.check(self.ctx._lib.pb2_comm_barrier(self.h))

    def peer_alloc(self, nbytes):
        """Collective: returns a ctypes array of nranks device pointers (peer buffers mapped in this process; [rank] = local)."""
        peers = (C.c_void_p * self.nranks)()
        self.ctx.check(self.ctx._lib.pb2_comm_peer_alloc(self.h, int(nbytes), peers))
        return peers

    def close(self):
        if self.h:
            self.ctx._lib.pb2_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


CONTACT_DTYPE = np.dtype([("point1", np.float32, (3,)), ("point2", np.float32, (3,)), ("normal1", np.float32, (3,)),
                          ("normal2", np.float32, (3,)), ("dist", np.float32)])
assert CONTACT_DTYPE.itemsize == 52


def contact(shapes, shape1, pos1, shape2, pos2, prediction, out=None):
    """query::contact(pos1, g1, pos2, g2, prediction), batched (contact_shape_shape.rs:123-138).
    Returns (contacts (n, 13) f32 [point1, point2, normal1, normal2, dist], status (n,) u8: 0 None, 1 Some,
    2 Unsupported)."""
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    if out is not None:  # caller-owned (e.g. page-locked) output buffers: (contacts (n, 13) f32, status (n,) u8)
        out, status = out
        po = out.data_ptr() if _is_torch(out) else out.ctypes.data
        pst = status.data_ptr() if _is_torch(status) else status.ctypes.data
    else:
        out, po = _empty((n, 13), np.float32, mem, ctx.torch_device)
        status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_contact_batch(ctx.h, shapes.h, ps1, ps2, p1, p2, float(prediction), n, po, pst, None, mem))
    return out, status


def contact_compact(shapes, shape1, pos1, shape2, pos2, prediction, capacity=None):
    """Same query, compacted output: (contacts (count, 13), pair_index (count,))."""
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    cap = int(capacity) if capacity is not None else n
    out, po = _empty((cap, 13), np.float32, mem, ctx.torch_device)
    idx, pi = _empty((cap,), np.uint32, mem, ctx.torch_device)
    cnt = C.c_uint64(0)
    ctx.check(ctx._lib.pb2_contact_batch_compact(ctx.h, shapes.h, ps1, ps2, p1, p2, float(prediction), n, po, pi, cap,
                                                 C.byref(cnt), mem))
    c = int(cnt.value)
    return out[:c], idx[:c]


def contact_local(shapes, shape1, shape2, pos12, prediction):
    """QueryDispatcher::contact(pos12, g1, g2, prediction), batched (query_dispatcher.rs:430-436): the relative pose is the input and
    the contact stays in the shapes' local frames. Returns (contacts (n, 13), status (n,))."""
    ctx = shapes.ctx
    n = int(pos12.shape[0])
    kp, pp, mem = _prep(pos12, np.float32)
    k1, p1, _ = _prep(shape1, np.uint32, mem)
    k2, p2, _ = _prep(shape2, np.uint32, mem)
    out, po = _empty((n, 13), np.float32, mem, ctx.torch_device)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_contact_batch_local(ctx.h, shapes.h, p1, p2, pp, float(prediction), n, po, pst, mem))
    return out, status


def contact_pairs_compact(shapes, collider_shape, collider_pose, pairs, prediction, capacity=None):
    """query::contact for every broad-phase pair (a, b) of `pairs` ((n, 2) collider indices, e.g. the output of
    Bvh.traverse_bvtt_single_tree): collider i is shape collider_shape[i] at pose collider_pose[i]. Compacted output:
    (contacts (count, 13), pair_index (count,)) with pair_index pointing into `pairs`."""
    ctx = shapes.ctx
    n = int(pairs.shape[0])
    nc = int(collider_pose.shape[0])
    kp, pp, mem = _prep(collider_pose, np.float32)
    ks, ps, _ = _prep(collider_shape, np.uint32, mem)
    kab, pab, _ = _prep(pairs, np.uint32, mem)
    cap = int(capacity) if capacity is not None else n
    out, po = _empty((cap, 13), np.float32, mem, ctx.torch_device)
    idx, pi = _empty((cap,), np.uint32, mem, ctx.torch_device)
    cnt = C.c_uint64(0)
    ctx.check(ctx._lib.pb2_contact_pairs_compact(ctx.h, shapes.h, ps, pp, nc, pab, n, float(prediction), po, pi, cap, C.byref(cnt), mem))
    c = int(cnt.value)
    return out[:c], idx[:c]


def _pair_query(fn_name, out_dtype, shapes, shape1, pos1, shape2, pos2):
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    out, po = _empty((n,), out_dtype, mem, ctx.torch_device)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(getattr(ctx._lib, fn_name)(ctx.h, shapes.h, ps1, ps2, p1, p2, n, po, pst, mem))
    return out, status


def distance(shapes, shape1, pos1, shape2, pos2):
    """query::distance(pos1, g1, pos2, g2), batched (query/distance/distance.rs:89-97): (dist (n,) f32, status (n,) u8: 0 Ok,
    2 Unsupported)."""
    return _pair_query("pb2_distance_batch", np.float32, shapes, shape1, pos1, shape2, pos2)


def intersection_test(shapes, shape1, pos1, shape2, pos2):
    """query::intersection_test(pos1, g1, pos2, g2), batched (intersection_test.rs:88-96): (hit (n,) u8, status (n,) u8)."""
    return _pair_query("pb2_intersection_test_batch", np.uint8, shapes, shape1, pos1, shape2, pos2)


class ShapeCastOptions:
    """query::ShapeCastOptions (shape_cast.rs:196-243), same defaults."""

    def __init__(self, max_time_of_impact=float(np.finfo(np.float32).max), target_distance=0.0, stop_at_penetration=True,
                 compute_impact_geometry_on_penetration=True):
        self.max_time_of_impact = float(max_time_of_impact)
        self.target_distance = float(target_distance)
        self.stop_at_penetration = bool(stop_at_penetration)
        self.compute_impact_geometry_on_penetration = bool(compute_impact_geometry_on_penetration)

    @classmethod
    def with_max_time_of_impact(cls, max_time_of_impact):
        return cls(max_time_of_impact=max_time_of_impact)


def cast_shapes(shapes, shape1, pos1, vel1, shape2, pos2, vel2, options=None):
    """query::cast_shapes(pos1, vel1, g1, pos2, vel2, g2, options), batched (shape_cast.rs:268-286). Returns (hits (n, 13) f32 =
    witness1, witness2, normal1, normal2 in the shapes' local frames and time_of_impact last; status (n,) u8: 0 None,
    1 Converged, 2 PenetratingOrWithinTargetDist, 3 unknown shape, 4 host fallback)."""
    o = options or ShapeCastOptions()
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    kv1, v1, _ = _prep(vel1, np.float32, mem)
    kv2, v2, _ = _prep(vel2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    out, po = _empty((n, 13), np.float32, mem, ctx.torch_device)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_cast_shapes_batch(ctx.h, shapes.h, ps1, ps2, p1, v1, p2, v2, o.max_time_of_impact, o.target_distance,
                                             int(o.stop_at_penetration), int(o.compute_impact_geometry_on_penetration), n, po, pst, mem))
    return out, status


def contact_manifolds(shapes, shape1, pos1, shape2, pos2, prediction, max_points=8):
    """QueryDispatcher::contact_manifolds(pos1.inv_mul(pos2), g1, g2, prediction, ..) on empty manifolds, batched: Ball / Cuboid
    pairs through the closed-form arms, pairs with a ConvexPolyhedron (vs Cuboid / ConvexPolyhedron) through pfm_pfm once
    Shapes.set_hull_topology was called. Returns (normals (n, 6) = local_n1, local_n2; counts (n,) u32; points (n, max_points, 9) f32 = local_p1,
    local_p2, dist, fid1, fid2 (the last two are PackedFeatureId bit patterns: view as u32); status (n,) u8: 0 ok,
    2 unsupported pair, 3 host fallback, 4 more than max_points contacts)."""
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    normals, pn = _empty((n, 6), np.float32, mem, ctx.torch_device)
    counts, pc = _empty((n,), np.uint32, mem, ctx.torch_device)
    points, pp = _empty((n, int(max_points), 9), np.float32, mem, ctx.torch_device)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_contact_manifolds_batch(ctx.h, shapes.h, ps1, ps2, p1, p2, float(prediction), n, int(max_points), pn, pc, pp, pst, mem))
    return normals, counts, points, status


def _inout(x, dtype, mem):
    """A private copy of an in / out array in the memory space of the call: (array, address)."""
    if mem == MEM_DEVICE:
        tdt = {np.float32: torch.float32, np.uint32: torch.int32}[dtype]
        t = (x if _is_torch(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=dtype).view(np.int32 if dtype is np.uint32 else dtype)))
        t = t.to(device=None if t.is_cuda else "cuda", dtype=tdt).contiguous().clone()
        # the copy is a kernel on torch's current stream; the library's stream is non-blocking and would not wait for it
        torch.cuda.current_stream(t.device).synchronize()
        return t, t.data_ptr()
    a = np.array(x.cpu().numpy() if _is_torch(x) else x, dtype=dtype, order="C", copy=True)
    return a, a.ctypes.data


def manifolds_try_update(ctx, pos1, pos2, normals, counts, points, angle_dot_threshold=0.99984769515, dist_sq_threshold=1.0e-6):
    """ContactManifold::try_update_contacts[_eps](pos1.inv_mul(pos2)) (contact_manifold.rs:652-699) on manifolds in the layout
    contact_manifolds returns. Returns (kept (n,) u8, points'): points' carries the refreshed dist / local_p1 (for a rejected
    manifold the points visited before the rejecting one, as the reference leaves them)."""
    n = int(pos1.shape[0])
    max_points = int(points.shape[1])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    kn, pn, _ = _prep(normals, np.float32, mem)
    kc, pc, _ = _prep(counts, np.uint32, mem)
    pts, pp = _inout(points, np.float32, mem)
    kept, pk = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_manifolds_try_update(ctx.h, p1, p2, n, max_points, float(angle_dot_threshold), float(dist_sq_threshold), pn, pc, pp, pk, mem))
    return kept, pts


def contact_manifolds_update(shapes, shape1, pos1, shape2, pos2, prediction, normals, counts, points, with_match=True):
    """QueryDispatcher::contact_manifolds called with last frame's manifolds (normals, counts, points as contact_manifolds returned
    them): the cuboid-cuboid and pfm_pfm arms keep a manifold that passes try_update_contacts, everything else is recomputed.
    Returns (normals, counts, points, status, kept (n,) u8, match (n, max_points) i32 or None): match[k, i] = index of last
    frame's point whose ContactData ContactManifold::match_contacts would hand to new point i (-1: none)."""
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    max_points = int(points.shape[1])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    nr, pn = _inout(normals, np.float32, mem)
    ct, pc = _inout(counts, np.uint32, mem)
    pts, pp = _inout(points, np.float32, mem)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    kept, pk = _empty((n,), np.uint8, mem, ctx.torch_device)
    match, pm = _empty((n, max_points), np.int32, mem, ctx.torch_device) if with_match else (None, None)
    ctx.check(ctx._lib.pb2_contact_manifolds_update_batch(ctx.h, shapes.h, ps1, ps2, p1, p2, float(prediction), n, max_points, pn, pc, pp, pst, pk,
                                                          pm, mem))
    return nr, ct, pts, status, kept, match


def closest_points(shapes, shape1, pos1, shape2, pos2, max_dist):
    """query::closest_points(pos1, g1, pos2, g2, max_dist), batched (closest_points_shape_shape.rs:220-231). Returns (points (n, 6)
    f32 = world-space p1, p2 (zeros unless WithinMargin); kind (n,) u8: 0 Disjoint, 1 WithinMargin, 2 Intersecting; status (n,)
    u8: 1 ok, 2 unknown shape, 3 host fallback)."""
    ctx = shapes.ctx
    n = int(pos1.shape[0])
    k1, p1, mem = _prep(pos1, np.float32)
    k2, p2, _ = _prep(pos2, np.float32, mem)
    ks1, ps1, _ = _prep(shape1, np.uint32, mem)
    ks2, ps2, _ = _prep(shape2, np.uint32, mem)
    out, po = _empty((n, 6), np.float32, mem, ctx.torch_device)
    kind, pk = _empty((n,), np.uint8, mem, ctx.torch_device)
    status, pst = _empty((n,), np.uint8, mem, ctx.torch_device)
    ctx.check(ctx._lib.pb2_closest_points_batch(ctx.h, shapes.h, ps1, ps2, p1, p2, float(max_dist), n, po, pk, pst, mem))
    return out, kind, status
