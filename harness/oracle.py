"""ctypes loader for the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by parry_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None
P, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()   # no-op unless a source under oracle/ is newer than the library
        l = C.CDLL(LIB_PATH)
        l.pb2o_hardware_threads.restype = i32
        l.pb2o_trimesh_create.restype = P
        l.pb2o_trimesh_create.argtypes = [P, u32, P, u32, i32]
        l.pb2o_trimesh_destroy.argtypes = [P]
        l.pb2o_trimesh_num_nodes.restype = u32
        l.pb2o_trimesh_num_nodes.argtypes = [P]
        l.pb2o_trimesh_copy_nodes.argtypes = [P, P]
        l.pb2o_trimesh_cast_rays.argtypes = [P, P, P, u32, f32, i32, i32, i32, P, P, P, P]
        l.pb2o_trimesh_contact_batch.argtypes = [P, P, P, P, P, P, P, f32, u32, i32, i32, P, P, P]
        l.pb2o_trimesh_contact_batch.restype = None
        l.pb2o_trimesh_distance_batch.argtypes = [P] * 7 + [u32, i32, i32, P, P]
        l.pb2o_trimesh_distance_batch.restype = None
        l.pb2o_compound_trimesh_contact_batch.argtypes = [P] * 11 + [f32, u32, i32, i32, i32, P, P, P]
        l.pb2o_compound_trimesh_contact_batch.restype = None
        l.pb2o_trimesh_cast_shapes.argtypes = [P, P, P, P, P, P, P, u32, P, P, i32, f32, f32, i32, i32, P, P]
        l.pb2o_trimesh_cast_shapes.restype = i32
        l.pb2o_trimesh_project_points.argtypes = [P, P, P, u32, i32, i32, i32, P, P, P]
        l.pb2o_trimesh_project_points.restype = None
        l.pb2o_bvh_create.restype = P
        l.pb2o_bvh_create.argtypes = [P, u32, i32]
        l.pb2o_bvh_destroy.argtypes = [P]
        l.pb2o_bvh_num_nodes.restype = u32
        l.pb2o_bvh_num_nodes.argtypes = [P]
        l.pb2o_bvh_copy_nodes.argtypes = [P, P]
        l.pb2o_bvh_copy_parents.argtypes = [P, P]
        l.pb2o_bvh_copy_leaf_node_indices.argtypes = [P, P]
        l.pb2o_bvh_update_leaves.argtypes = [P, P, P, u32, f32]
        l.pb2o_bvh_refit.argtypes = [P]
        l.pb2o_bvh_refit_without_opt.argtypes = [P]
        l.pb2o_bvh_rebuild.argtypes = [P, C.c_int]
        l.pb2o_bvh_intersect_aabbs.restype = u64
        l.pb2o_bvh_intersect_aabbs.argtypes = [P, P, u32, i32, P, P, u64]
        l.pb2o_bvh_self_pairs.restype = u64
        l.pb2o_bvh_self_pairs.argtypes = [P, i32, P, u64]
        l.pb2o_bvh_leaf_pairs.restype = u64
        l.pb2o_bvh_leaf_pairs.argtypes = [P, P, P, u64]
        l.pb2o_bvh_cast_rays_shapes.argtypes = [P, P, P, P, P, u32, f32, i32, i32, P, P, P, P]
        l.pb2o_bvh_project_points_shapes.argtypes = [P, P, P, P, P, P, P, P, u32, f32, i32, i32, P, P, P]
        l.pb2o_bvh_project_points_shapes.restype = None
        l.pb2o_bvh_cast_rays_shapes2.argtypes = [P, P, P, P, P, P, P, P, u32, f32, i32, i32, P, P, P, P]
        l.pb2o_cast_shapes_batch.argtypes = [P, P, P, P, P, P, P, P, P, f32, f32, i32, i32, u32, i32, P, P]
        l.pb2o_compound_contact_batch.argtypes = [P, P, P, P, P, P, P, P, P, P, P, f32, i32, u32, i32, P, P, P]
        l.pb2o_contact_manifolds_batch.argtypes = [P, P, P, P, P, P, P, f32, u32, u32, i32, P, P, P, P]
        l.pb2o_contact_manifolds_batch2.argtypes = [P] * 20 + [f32, u32, u32, i32, P, P, P, P]
        l.pb2o_closest_points_batch.argtypes = [P, P, P, P, P, P, P, f32, u32, i32, P, P, P]
        l.pb2o_manifolds_try_update.argtypes = [P, P, u32, u32, P, P, P, P]
        l.pb2o_contact_local_batch.argtypes = [P, P, P, P, P, P, P, f32, u32, i32, P, P]
        l.pb2o_contact_manifolds_update_batch.argtypes = [P] * 20 + [f32, u32, u32, i32, i32, P, P, P, P, P, P]
        l.pb2o_compound_compound_contact_batch.argtypes = [P, P, P, P, P, P, P, P, P, P, P, f32, u32, i32, P, P, P]
        l.pb2o_convex_cast_ray.restype = i32
        l.pb2o_convex_cast_ray.argtypes = [P, u32, P, P, f32, i32, P, P]
        l.pb2o_shape_cast_ray.restype = i32
        l.pb2o_shape_cast_ray.argtypes = [i32, P, P, P, f32, i32, P, P, P]
        l.pb2o_shape_cast_ray_toi.restype = i32
        l.pb2o_shape_cast_ray_toi.argtypes = [i32, P, P, P, f32, i32, P]
        for name, res, args in _EXTRA:
            fn = getattr(l, name, None)
            if fn is not None:
                fn.restype = res
                fn.argtypes = args
        _lib = l
    return _lib


_EXTRA = []  # (name, restype, argtypes) registered by later sections


def hardware_threads():
    return int(lib().pb2o_hardware_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


NODE_HALF = np.dtype([("mins", np.float32, (3,)), ("children", np.uint32), ("maxs", np.float32, (3,)), ("data", np.uint32)])
NODE_WIDE = np.dtype([("left", NODE_HALF), ("right", NODE_HALF)])


class TriMesh:
    def __init__(self, vertices, indices, strategy=0):
        self.v, self.i = _f32(vertices), _u32(indices)
        self.nt = self.i.shape[0]
        self.h = lib().pb2o_trimesh_create(self.v.ctypes.data, self.v.shape[0], self.i.ctypes.data, self.nt, strategy)

    def nodes(self):
        n = lib().pb2o_trimesh_num_nodes(self.h)
        out = np.zeros(n, dtype=NODE_WIDE)
        lib().pb2o_trimesh_copy_nodes(self.h, out.ctypes.data)
        return out

    def cast_rays(self, pose, rays, max_toi, solid=True, with_normal=False, mode=0, threads=1):
        rays = _f32(rays)
        m = rays.shape[0]
        pose = None if pose is None else _f32(pose)
        toi = np.zeros(m, dtype=np.float32)
        tri = np.zeros(m, dtype=np.uint32)
        normal = np.zeros((m, 3), dtype=np.float32) if with_normal else None
        feature = np.zeros(m, dtype=np.uint32) if with_normal else None
        lib().pb2o_trimesh_cast_rays(self.h, None if pose is None else pose.ctypes.data, rays.ctypes.data, m, max_toi, int(solid),
                                     mode, threads, toi.ctypes.data, tri.ctypes.data,
                                     None if normal is None else normal.ctypes.data,
                                     None if feature is None else feature.ctypes.data)
        return (toi, tri, normal, feature) if with_normal else (toi, tri)

    def project_points(self, pose, points, solid=True, mode=0, threads=1):
        """PointQuery::project_point on the mesh: (proj (n,3), inside (n,) u8, tri (n,) u32). mode 1 = brute force, min-index ties."""
        pts = _f32(points)
        n = pts.shape[0]
        pose = None if pose is None else _f32(pose)
        proj = np.zeros((n, 3), dtype=np.float32)
        inside = np.zeros(n, dtype=np.uint8)
        tri = np.zeros(n, dtype=np.uint32)
        lib().pb2o_trimesh_project_points(self.h, None if pose is None else pose.ctypes.data, pts.ctypes.data, n, int(solid), mode, threads,
                                          proj.ctypes.data, inside.ctypes.data, tri.ctypes.data)
        return proj, inside, tri

    def contact_shapes(self, mesh_pose, table, shape2, pos2, prediction, threads=1, min_index_ties=False):
        """query::contact(mesh_pose, mesh, pos2[k], shape2[k], prediction): (contacts (n,13), status, part)."""
        s2, p2, mp = _u32(shape2), _f32(pos2), _f32(mesh_pose)
        n = len(s2)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        part = np.zeros(n, dtype=np.uint32)
        lib().pb2o_trimesh_contact_batch(self.h, mp.ctypes.data, table.kinds.ctypes.data, table.params.ctypes.data, table.points.ctypes.data,
                                         s2.ctypes.data, p2.ctypes.data, prediction, n, threads, int(min_index_ties), out.ctypes.data,
                                         status.ctypes.data, part.ctypes.data)
        return out, status, part

    def distance_shapes(self, mesh_pose, table, shape_ids, poses, mesh_second=False, threads=1):
        """query::distance(mesh_pose, mesh, poses[k], shape k) (or with the shape first): (dist (n,), part (n,) closest triangle)."""
        s2, p2, mp = _u32(shape_ids), _f32(poses), _f32(mesh_pose)
        n = len(s2)
        dist = np.zeros(n, dtype=np.float32)
        part = np.zeros(n, dtype=np.uint32)
        lib().pb2o_trimesh_distance_batch(self.h, mp.ctypes.data, table.kinds.ctypes.data, table.params.ctypes.data, table.points.ctypes.data,
                                          s2.ctypes.data, p2.ctypes.data, n, threads, int(mesh_second), dist.ctypes.data, part.ctypes.data)
        return dist, part

    def contact_compounds(self, mesh_pose, table, comp_first, comp_count, part_shape, part_pose, ids, poses, prediction, trimesh_first=False,
                          threads=1, min_index_ties=False):
        """query::contact between Compound ids[k] at poses[k] and this mesh at mesh_pose (oracle groundwork, no GPU path yet);
        trimesh_first: the mesh is shape 1. Returns (contacts (n,13), status, parts (n,2) = winning {compound part, triangle})."""
        cf, cc, psid, a = _u32(comp_first), _u32(comp_count), _u32(part_shape), _u32(ids)
        pp, p, mp = _f32(part_pose), _f32(poses), _f32(mesh_pose)
        n = len(a)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        parts = np.zeros((n, 2), dtype=np.uint32)
        lib().pb2o_compound_trimesh_contact_batch(self.h, mp.ctypes.data, table.kinds.ctypes.data, table.params.ctypes.data, table.points.ctypes.data,
                                                  cf.ctypes.data, cc.ctypes.data, psid.ctypes.data, pp.ctypes.data, a.ctypes.data, p.ctypes.data,
                                                  prediction, n, threads, int(trimesh_first), int(min_index_ties), out.ctypes.data,
                                                  status.ctypes.data, parts.ctypes.data)
        return out, status, parts

    def cast_shapes(self, mesh_pose, mesh_vel, pose, vel, other_mesh=None, table=None, shape=0, mesh_second=False, max_toi=None,
                    target_distance=0.0, stop_at_penetration=True, compute_impact_geometry_on_penetration=True):
        """query::cast_shapes with this mesh as shape 1 (or 2 with mesh_second) against another TriMesh or table.shape (oracle
        groundwork, no GPU path yet). Returns None or (out (13,), status, part)."""
        mp, mv, po, ve = _f32(mesh_pose), _f32(mesh_vel), _f32(pose), _f32(vel)
        out = np.zeros(13, dtype=np.float32)
        part = C.c_uint32(0)
        tk = table.kinds.ctypes.data if table is not None else None
        tp = table.params.ctypes.data if table is not None else None
        tq = table.points.ctypes.data if table is not None else None
        st = lib().pb2o_trimesh_cast_shapes(self.h, mp.ctypes.data, mv.ctypes.data, other_mesh.h if other_mesh is not None else None, tk, tp, tq,
                                            int(shape), po.ctypes.data, ve.ctypes.data, int(mesh_second),
                                            float(np.finfo(np.float32).max) if max_toi is None else max_toi, target_distance,
                                            int(stop_at_penetration), int(compute_impact_geometry_on_penetration), out.ctypes.data, C.byref(part))
        return None if st == 0 else (out, st, part.value)

    def __del__(self):
        try:
            lib().pb2o_trimesh_destroy(self.h)
        except Exception:
            pass


class Bvh:
    def __init__(self, aabbs, strategy=0):
        self.aabbs = _f32(aabbs).reshape(-1, 6)
        self.n = self.aabbs.shape[0]
        self.h = lib().pb2o_bvh_create(self.aabbs.ctypes.data, self.n, strategy)

    def nodes(self):
        n = lib().pb2o_bvh_num_nodes(self.h)
        out = np.zeros(n, dtype=NODE_WIDE)
        if n:
            lib().pb2o_bvh_copy_nodes(self.h, out.ctypes.data)
        return out

    def parents(self):
        n = lib().pb2o_bvh_num_nodes(self.h)
        out = np.zeros(n, dtype=np.uint64)
        if n:
            lib().pb2o_bvh_copy_parents(self.h, out.ctypes.data)
        return out

    def leaf_node_indices(self):
        out = np.zeros(self.n, dtype=np.uint64)
        if self.n:
            lib().pb2o_bvh_copy_leaf_node_indices(self.h, out.ctypes.data)
        return out

    def update_leaves(self, aabbs, ids=None, margin=0.0):
        a = _f32(aabbs)
        i = None if ids is None else _u32(ids)
        lib().pb2o_bvh_update_leaves(self.h, None if i is None else i.ctypes.data, a.ctypes.data, a.shape[0], margin)

    def refit(self):
        lib().pb2o_bvh_refit(self.h)

    def refit_without_opt(self):
        lib().pb2o_bvh_refit_without_opt(self.h)

    def rebuild(self, strategy=0):
        """Bvh::rebuild (bvh_binned_build.rs:11-36): leaves keep their change flags, nothing is resolved."""
        lib().pb2o_bvh_rebuild(self.h, int(strategy))

    def intersect_aabbs(self, queries, threads=1):
        q = _f32(queries).reshape(-1, 6)
        m = q.shape[0]
        offs = np.zeros(m + 1, dtype=np.uint32)
        total = lib().pb2o_bvh_intersect_aabbs(self.h, q.ctypes.data, m, threads, offs.ctypes.data, None, 0)
        ids = np.zeros(max(1, total), dtype=np.uint32)
        lib().pb2o_bvh_intersect_aabbs(self.h, q.ctypes.data, m, threads, offs.ctypes.data, ids.ctypes.data, total)
        return offs, ids[:total]

    def self_pairs(self, change_detection=False):
        total = lib().pb2o_bvh_self_pairs(self.h, int(change_detection), None, 0)
        pairs = np.zeros((max(1, total), 2), dtype=np.uint32)
        lib().pb2o_bvh_self_pairs(self.h, int(change_detection), pairs.ctypes.data, total)
        return pairs[:total]

    def leaf_pairs(self, other):
        total = lib().pb2o_bvh_leaf_pairs(self.h, other.h, None, 0)
        pairs = np.zeros((max(1, total), 2), dtype=np.uint32)
        lib().pb2o_bvh_leaf_pairs(self.h, other.h, pairs.ctypes.data, total)
        return pairs[:total]

    def cast_rays_shapes(self, kinds, params, poses, rays, max_toi, solid=True, with_normal=False, threads=1, points=None, first=None,
                         count=None):
        """Leaf i = shape kinds[i] (0 ball, 1 cuboid, 2 ConvexPolyhedron = points[first[i] : first[i] + count[i]]) at poses[i]."""
        rays = _f32(rays)
        m = rays.shape[0]
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        params = _f32(params).reshape(-1, 3)
        poses = _f32(poses)
        toi = np.zeros(m, dtype=np.float32)
        leaf = np.zeros(m, dtype=np.uint32)
        normal = np.zeros((m, 3), dtype=np.float32) if with_normal else None
        feature = np.zeros(m, dtype=np.uint32) if with_normal else None
        if points is not None:
            points = _f32(points).reshape(-1, 3)
            first = np.ascontiguousarray(first, dtype=np.uint32)
            count = np.ascontiguousarray(count, dtype=np.uint32)
        else:
            assert not (kinds == 2).any()
        lib().pb2o_bvh_cast_rays_shapes2(self.h, kinds.ctypes.data, params.ctypes.data,
                                         None if points is None else points.ctypes.data, None if points is None else first.ctypes.data,
                                         None if points is None else count.ctypes.data, poses.ctypes.data, rays.ctypes.data, m, max_toi,
                                         int(solid), threads, toi.ctypes.data, leaf.ctypes.data,
                                         None if normal is None else normal.ctypes.data,
                                         None if feature is None else feature.ctypes.data)
        return (toi, leaf, normal, feature) if with_normal else (toi, leaf)

    def project_points_shapes(self, kinds, params, poses, pts, max_distance, solid=True, threads=1, points=None, first=None, count=None):
        """Bvh::project_point with typed leaves (as cast_rays_shapes): (proj (m,3) world space, inside (m,), leaf (m,))."""
        pts = _f32(pts)
        m = pts.shape[0]
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        params = _f32(params).reshape(-1, 3)
        poses = _f32(poses)
        proj = np.zeros((m, 3), dtype=np.float32)
        inside = np.zeros(m, dtype=np.uint8)
        leaf = np.zeros(m, dtype=np.uint32)
        if points is not None:
            points = _f32(points).reshape(-1, 3)
            first = np.ascontiguousarray(first, dtype=np.uint32)
            count = np.ascontiguousarray(count, dtype=np.uint32)
        else:
            assert not (kinds == 2).any()
        lib().pb2o_bvh_project_points_shapes(self.h, kinds.ctypes.data, params.ctypes.data, None if points is None else points.ctypes.data,
                                             None if points is None else first.ctypes.data, None if points is None else count.ctypes.data,
                                             poses.ctypes.data, pts.ctypes.data, m, max_distance, int(solid), threads, proj.ctypes.data,
                                             inside.ctypes.data, leaf.ctypes.data)
        return proj, inside, leaf

    def __del__(self):
        try:
            lib().pb2o_bvh_destroy(self.h)
        except Exception:
            pass


def convex_cast_ray(points, pose, ray, max_toi, solid=True):
    """RayCast::cast_ray_and_get_normal for one ConvexPolyhedron (ray_support_map.rs:163-181). Returns None or (toi, normal)."""
    p = _f32(points).reshape(-1, 3)
    po = None if pose is None else _f32(pose)
    r = _f32(ray)
    toi = C.c_float(0.0)
    n = np.zeros(3, dtype=np.float32)
    hit = lib().pb2o_convex_cast_ray(p.ctypes.data, p.shape[0], None if po is None else po.ctypes.data, r.ctypes.data, max_toi, int(solid),
                                     C.byref(toi), n.ctypes.data)
    return (np.float32(toi.value), n) if hit else None


def shape_cast_ray(kind, params, pose, ray, max_toi, solid=True):
    """RayCast::cast_ray_and_get_normal for a single Ball(0)/Cuboid(1)/Triangle(2). Returns None or (toi, normal, feature)."""
    p = _f32(params).ravel()
    r = _f32(ray).ravel()
    po = None if pose is None else _f32(pose).ravel()
    toi = C.c_float(0)
    n = np.zeros(3, dtype=np.float32)
    f = C.c_uint32(0)
    hit = lib().pb2o_shape_cast_ray(kind, p.ctypes.data, None if po is None else po.ctypes.data, r.ctypes.data, max_toi, int(solid),
                                    C.addressof(toi), n.ctypes.data, C.addressof(f))
    return (toi.value, n, f.value) if hit else None


def shape_cast_ray_toi(kind, params, pose, ray, max_toi, solid=True):
    p = _f32(params).ravel()
    r = _f32(ray).ravel()
    po = None if pose is None else _f32(pose).ravel()
    toi = C.c_float(0)
    hit = lib().pb2o_shape_cast_ray_toi(kind, p.ctypes.data, None if po is None else po.ctypes.data, r.ctypes.data, max_toi, int(solid),
                                        C.addressof(toi))
    return toi.value if hit else None


_EXTRA.append(("pb2o_clip_aabb_line", i32, [P, P, P, P]))


def clip_aabb_line(mins, maxs, origin, direction):
    """query::details::clip_aabb_line (clip_aabb_line.rs:79-187). Returns None or (near t, far t)."""
    box = _f32(list(mins) + list(maxs)).ravel()
    o, d = _f32(origin).ravel(), _f32(direction).ravel()
    nf = np.zeros(2, dtype=np.float32)
    hit = lib().pb2o_clip_aabb_line(box.ctypes.data, o.ctypes.data, d.ctypes.data, nf.ctypes.data)
    return (nf[0], nf[1]) if hit else None


_EXTRA.append(("pb2o_shape_aabbs", None, [P, P, P, P, P, P, u32, P]))


def shape_aabbs(kinds, params, poses, points=None, first=None, count=None):
    """Shape::compute_aabb(pos) per collider. params: (n,3); convex: points (np,3), first/count per collider."""
    kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
    params = _f32(params).reshape(-1, 3)
    poses = _f32(poses)
    n = len(kinds)
    out = np.zeros((n, 6), dtype=np.float32)
    pts = None if points is None else _f32(points)
    fi = None if first is None else _u32(first)
    ct = None if count is None else _u32(count)
    lib().pb2o_shape_aabbs(kinds.ctypes.data, params.ctypes.data, None if pts is None else pts.ctypes.data,
                           None if fi is None else fi.ctypes.data, None if ct is None else ct.ctypes.data, poses.ctypes.data, n,
                           out.ctypes.data)
    return out


# ---------------------------------------------------------------- query::contact
_EXTRA.append(("pb2o_contact_batch", None, [P, P, P, P, P, P, P, f32, u32, i32, P, P, P]))
_EXTRA.append(("pb2o_dispatch_contact", i32, [P, P, P, u32, u32, P, f32, P]))
_EXTRA.append(("pb2o_distance_batch", None, [P, P, P, u32, P, P, P, P, u32, i32, P, P]))
_EXTRA.append(("pb2o_intersection_test_batch", None, [P, P, P, u32, P, P, P, P, u32, i32, P, P]))
_EXTRA.append(("pb2o_gjk_closest_points", i32, [P, P, P, u32, u32, P, f32, P]))


class ShapeTable:
    """Same table layout as pb2_shapes_create: kinds[n], params[n,4], points[np,3]. shapes: list of ('ball', r) /
    ('cuboid', he) / ('convex', pts)."""

    def __init__(self, shapes):
        n = len(shapes)
        self.kinds = np.zeros(n, dtype=np.uint8)
        self.params = np.zeros((n, 4), dtype=np.float32)
        pu = self.params.view(np.uint32)
        pts, npts = [], 0
        for i, (k, v) in enumerate(shapes):
            if k == "ball":
                self.kinds[i] = 0
                self.params[i, 0] = v
            elif k == "cuboid":
                self.kinds[i] = 1
                self.params[i, :3] = v
            else:
                self.kinds[i] = 3 if k == "triangle" else 2
                p = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
                pu[i, 0], pu[i, 1] = npts, len(p)
                pts.append(p)
                npts += len(p)
        self.points = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((1, 3), np.float32), dtype=np.float32)

    def contact(self, shape1, pos1, shape2, pos2, prediction, threads=1, with_stats=False):
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        stats = np.zeros((n, 6), dtype=np.int32) if with_stats else None
        lib().pb2o_contact_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1.ctypes.data, s2.ctypes.data,
                                 p1.ctypes.data, p2.ctypes.data, prediction, n, threads, out.ctypes.data, status.ctypes.data,
                                 None if stats is None else stats.ctypes.data)
        return (out, status, stats) if with_stats else (out, status)

    def contact_local(self, shape1, pos1, shape2, pos2, prediction, threads=1):
        """QueryDispatcher::contact(pos1.inv_mul(pos2), ..) per pair, result left in the shapes' local frames: (out (n,13), status)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        lib().pb2o_contact_local_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1.ctypes.data, s2.ctypes.data,
                                       p1.ctypes.data, p2.ctypes.data, prediction, n, threads, out.ctypes.data, status.ctypes.data)
        return out, status

    def cast_shapes(self, shape1, pos1, vel1, shape2, pos2, vel2, max_time_of_impact=float(np.finfo(np.float32).max), target_distance=0.0,
                    stop_at_penetration=True, compute_impact_geometry_on_penetration=True, threads=1):
        """query::cast_shapes per pair: (out (n,13) = witness1, witness2, normal1, normal2 (local frames), toi; status: 0 None,
        1 Converged, 2 PenetratingOrWithinTargetDist)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        v1, v2 = _f32(vel1).reshape(-1, 3), _f32(vel2).reshape(-1, 3)
        n = len(s1)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        lib().pb2o_cast_shapes_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1.ctypes.data, s2.ctypes.data,
                                     p1.ctypes.data, v1.ctypes.data, p2.ctypes.data, v2.ctypes.data, max_time_of_impact, target_distance,
                                     int(stop_at_penetration), int(compute_impact_geometry_on_penetration), n, threads, out.ctypes.data,
                                     status.ctypes.data)
        return out, status

    def contact_compound(self, comp_first, comp_count, part_shape, part_pose, compound_id, pos_c, shape, pos_s, prediction,
                         compound_second=False, threads=1):
        """query::contact(pos_c, Compound, pos_s, shape) per pair (or the flipped call when compound_second): (out (n,13), status,
        part). Compound c = parts comp_first[c] .. + comp_count[c] of (part_shape -> this table, part_pose)."""
        cf, cc, psid, cid, sid = _u32(comp_first), _u32(comp_count), _u32(part_shape), _u32(compound_id), _u32(shape)
        pp, pc, ps = _f32(part_pose), _f32(pos_c), _f32(pos_s)
        n = len(cid)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        part = np.zeros(n, dtype=np.uint32)
        lib().pb2o_compound_contact_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, cf.ctypes.data, cc.ctypes.data,
                                          psid.ctypes.data, pp.ctypes.data, cid.ctypes.data, pc.ctypes.data, sid.ctypes.data, ps.ctypes.data,
                                          prediction, int(compound_second), n, threads, out.ctypes.data, status.ctypes.data, part.ctypes.data)
        return out, status, part

    @staticmethod
    def manifolds_try_update(pos1, pos2, normals, counts, points):
        """ContactManifold::try_update_contacts on manifolds as returned by contact_manifolds; returns (kept (n,) u8, points') with
        dists / local_p1 refreshed for the points visited before a rejection (in-place semantics of the reference)."""
        p1, p2 = _f32(pos1), _f32(pos2)
        n, mp = len(counts), points.shape[1]
        nr = np.ascontiguousarray(normals, dtype=np.float32).copy()
        ct = np.ascontiguousarray(counts, dtype=np.uint32)
        pts = np.ascontiguousarray(points, dtype=np.float32).copy()
        kept = np.zeros(n, dtype=np.uint8)
        lib().pb2o_manifolds_try_update(p1.ctypes.data, p2.ctypes.data, n, mp, nr.ctypes.data, ct.ctypes.data, pts.ctypes.data, kept.ctypes.data)
        return kept, pts

    def hull_topology(self):
        """Face topology of every ConvexPolyhedron of the table (harness/hull_topology.py), indexed per table entry:
        dict(hull_face_first, hull_face_count (both len(kinds)), face_normal, face_first, face_count, vertices_adj_to_face,
        edges_adj_to_face). Cached."""
        if getattr(self, "_topo", None) is None:
            from harness import hull_topology as ht
            pu = self.params.view(np.uint32)
            conv = np.nonzero(self.kinds == 2)[0]
            t = ht.hull_table([self.points[pu[i, 0]:pu[i, 0] + pu[i, 1]] for i in conv]) if len(conv) else None
            hf, hc = np.zeros(len(self.kinds), np.uint32), np.zeros(len(self.kinds), np.uint32)
            if t is not None:
                hf[conv], hc[conv] = t["hull_face_first"], t["hull_face_count"]
                t = dict(t)
                t["hull_face_first"], t["hull_face_count"] = hf, hc
                # vertex-side arrays are indexed by the table's global point index, the edge offset per table entry
                npnt = len(self.points)
                vfirst, vcount = np.zeros(npnt, np.uint32), np.zeros(npnt, np.uint32)
                hef = np.zeros(len(self.kinds), np.uint32)
                at = 0
                for j, i in enumerate(conv):
                    f0, c = int(pu[i, 0]), int(pu[i, 1])
                    vfirst[f0:f0 + c] = t["vert_first"][at:at + c]
                    vcount[f0:f0 + c] = t["vert_count"][at:at + c]
                    at += c
                    hef[i] = t["hull_edge_first"][j]
                t["vert_first"], t["vert_count"], t["hull_edge_first"] = vfirst, vcount, hef
            self._topo = t
        return self._topo

    def contact_manifolds(self, shape1, pos1, shape2, pos2, prediction, max_points=16, threads=1, topology=None):
        """contact_manifolds per pair, first frame: (normals (n,6), counts (n,), points (n,max_points,9) f32 with fid1/fid2 as u32 bit
        patterns in the last two columns, status (n,): 0 ok, 2 unsupported pair, 4 more than max_points)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        normals = np.zeros((n, 6), dtype=np.float32)
        counts = np.zeros(n, dtype=np.uint32)
        pts = np.zeros((n, max_points, 9), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        t = topology
        keys = ("hull_face_first", "hull_face_count", "face_normal", "face_first", "face_count", "vertices_adj_to_face", "edges_adj_to_face",
                "vert_first", "vert_count", "faces_adj_to_vertex", "edges_adj_to_vertex", "hull_edge_first", "edge_dir")
        keep = None if t is None else [np.ascontiguousarray(t[k]) if k in t else None for k in keys]
        tp = [None] * 13 if t is None else [None if a is None else a.ctypes.data for a in keep]
        lib().pb2o_contact_manifolds_batch2(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, *tp, s1.ctypes.data,
                                            s2.ctypes.data, p1.ctypes.data, p2.ctypes.data, prediction, n, max_points, threads,
                                            normals.ctypes.data, counts.ctypes.data, pts.ctypes.data, status.ctypes.data)
        return normals, counts, pts, status

    def contact_manifolds_update(self, shape1, pos1, shape2, pos2, prediction, normals, counts, points, threads=1, topology=None,
                                 seed_gjk=False):
        """contact_manifolds called again with last frame's manifolds (as returned by contact_manifolds): (normals, counts, points,
        status, kept (n,) u8, match (n,max_points) i32). seed_gjk: the pfm_pfm recomputation starts GJK from last frame's normal
        like the reference (contact_manifolds_pfm_pfm.rs:66); False restates the GPU path (default first direction)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n, max_points = len(s1), points.shape[1]
        normals = np.ascontiguousarray(normals, dtype=np.float32).copy()
        counts = np.ascontiguousarray(counts, dtype=np.uint32).copy()
        pts = np.ascontiguousarray(points, dtype=np.float32).copy()
        status = np.zeros(n, dtype=np.uint8)
        kept = np.zeros(n, dtype=np.uint8)
        match = np.full((n, max_points), -1, dtype=np.int32)
        t = topology
        keys = ("hull_face_first", "hull_face_count", "face_normal", "face_first", "face_count", "vertices_adj_to_face", "edges_adj_to_face",
                "vert_first", "vert_count", "faces_adj_to_vertex", "edges_adj_to_vertex", "hull_edge_first", "edge_dir")
        keep = None if t is None else [np.ascontiguousarray(t[k]) if k in t else None for k in keys]
        tp = [None] * 13 if t is None else [None if a is None else a.ctypes.data for a in keep]
        lib().pb2o_contact_manifolds_update_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, *tp, s1.ctypes.data,
                                                  s2.ctypes.data, p1.ctypes.data, p2.ctypes.data, prediction, n, max_points, threads,
                                                  int(bool(seed_gjk)), normals.ctypes.data, counts.ctypes.data, pts.ctypes.data,
                                                  status.ctypes.data, kept.ctypes.data, match.ctypes.data)
        return normals, counts, pts, status, kept, match

    def closest_points(self, shape1, pos1, shape2, pos2, max_dist, threads=1):
        """query::closest_points per pair: (points (n,6) world-space p1, p2; kind (n,): 0 Disjoint, 1 WithinMargin, 2 Intersecting;
        status (n,): 1 ok, 3 needs hull topology)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        out = np.zeros((n, 6), dtype=np.float32)
        kind = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint8)
        lib().pb2o_closest_points_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1.ctypes.data, s2.ctypes.data,
                                        p1.ctypes.data, p2.ctypes.data, max_dist, n, threads, out.ctypes.data, kind.ctypes.data, status.ctypes.data)
        return out, kind, status

    def contact_compound_compound(self, comp_first, comp_count, part_shape, part_pose, id1, pos1, id2, pos2, prediction, threads=1):
        """query::contact(pos1, Compound id1, pos2, Compound id2) per pair (oracle groundwork; no GPU path yet): (out (n,13),
        status, parts (n,2) winning part of each compound)."""
        cf, cc, psid, a, b = _u32(comp_first), _u32(comp_count), _u32(part_shape), _u32(id1), _u32(id2)
        pp, p1, p2 = _f32(part_pose), _f32(pos1), _f32(pos2)
        n = len(a)
        out = np.zeros((n, 13), dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        parts = np.zeros((n, 2), dtype=np.uint32)
        lib().pb2o_compound_compound_contact_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, cf.ctypes.data,
                                                   cc.ctypes.data, psid.ctypes.data, pp.ctypes.data, a.ctypes.data, p1.ctypes.data, b.ctypes.data,
                                                   p2.ctypes.data, prediction, n, threads, out.ctypes.data, status.ctypes.data, parts.ctypes.data)
        return out, status, parts

    def distance(self, shape1, pos1, shape2, pos2, threads=1):
        """query::distance per pair: (dist (n,), status (n,): 0 Ok, 2 Unsupported, 3 cuboid-cuboid)."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        out = np.zeros(n, dtype=np.float32)
        status = np.zeros(n, dtype=np.uint8)
        lib().pb2o_distance_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, len(self.kinds), s1.ctypes.data,
                                  s2.ctypes.data, p1.ctypes.data, p2.ctypes.data, n, threads, out.ctypes.data, status.ctypes.data)
        return out, status

    def intersection_test(self, shape1, pos1, shape2, pos2, threads=1):
        """query::intersection_test per pair: (hit (n,) u8, status (n,))."""
        s1, s2, p1, p2 = _u32(shape1), _u32(shape2), _f32(pos1), _f32(pos2)
        n = len(s1)
        out = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint8)
        lib().pb2o_intersection_test_batch(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, len(self.kinds),
                                           s1.ctypes.data, s2.ctypes.data, p1.ctypes.data, p2.ctypes.data, n, threads, out.ctypes.data,
                                           status.ctypes.data)
        return out, status

    def dispatch_contact(self, s1, s2, pos12, prediction):
        """DefaultQueryDispatcher::contact(pos12, g1, g2, prediction): (status, contact[13]) in local frames."""
        out = np.zeros(13, dtype=np.float32)
        p = _f32(pos12)
        st = lib().pb2o_dispatch_contact(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1, s2, p.ctypes.data,
                                         prediction, out.ctypes.data)
        return st, out

    def gjk_closest_points(self, s1, s2, pos12, max_dist):
        out = np.zeros(9, dtype=np.float32)
        p = _f32(pos12)
        kind = lib().pb2o_gjk_closest_points(self.kinds.ctypes.data, self.params.ctypes.data, self.points.ctypes.data, s1, s2, p.ctypes.data,
                                             max_dist, out.ctypes.data)
        return kind, out
