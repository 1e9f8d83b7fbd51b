//! parry-b200 — parry3d's query hot path (`Bvh`, `RayCast for TriMesh`, `QueryDispatcher`) served by the B200 CUDA library
//! behind `include/parry_b200.h`. Same names, argument meaning and error behaviour as the reference; the differences are the
//! ones the batched ABI forces and are spelled out on each item:
//!
//!  * one query per call becomes one *batch* per call (slices in, `Vec`s out): a GPU launch per ray or per pair would be all
//!    latency. Closures cannot cross to the GPU, so `Bvh::cast_ray`'s leaf callback becomes a typed shape table.
//!  * `QueryDispatcher` is a per-pair trait. [`B200Dispatcher`] answers it from a batch the caller *registers* before the
//!    per-pair calls start (`prepare_contacts`, …): rapier's narrow phase knows every pair of a step up front. A pair that
//!    was not registered is run as a batch of one; a pair with a shape the tables do not know returns `Err(Unsupported)`, so
//!    that `B200Dispatcher::new(..).chain(DefaultQueryDispatcher)` (query_dispatcher.rs:473) falls through to the CPU.
//!
//! NOT COMPILED in the authoring image (no cargo / rustc there); written against parry3d 0.25 as checked out under
//! /root/reference. File:line citations are relative to that checkout's `src/`.
use core::ffi::c_void;
use std::collections::HashMap;
use std::sync::Mutex;

use parry3d::bounding_volume::Aabb;
use parry3d::math::{Isometry, Point, Real, Vector};
use parry3d::partitioning::{BvhBuildStrategy, BvhNodeWide};
use parry3d::query::{
    ClosestPoints, Contact, NonlinearRigidMotion, QueryDispatcher, Ray, RayIntersection, ShapeCastHit, ShapeCastOptions,
    ShapeCastStatus, Unsupported,
};
use parry3d::shape::{FeatureId, Shape, TypedShape};
use parry_b200_sys as sys;

// ------------------------------------------------------------------------------------------------ errors / context
/// What a `pb2_status` other than `PB2_OK` becomes. `Overflow(required)` is retried inside the shim and never reaches callers.
#[derive(Debug, Clone, PartialEq)]
pub enum Error {
    Invalid(String),
    Cuda(String),
    Overflow(u64),
    Unsupported,
    /// `PB2_ERR_DEPTH`: a tree walk ran out of its fixed stack (cannot happen for Morton-linked trees).
    Depth(String),
}

/// One CUDA device + one stream (`pb2_ctx`). The C context is single-stream and not re-entrant — the same contract as a
/// `&mut BvhWorkspace` in the reference (bvh_tree.rs:135-145) — so every call goes through the mutex; distinct `B200`s are
/// independent. `Send + Sync` because `QueryDispatcher: Send + Sync` (query_dispatcher.rs:408).
pub struct B200 {
    ctx: Mutex<*mut sys::pb2_ctx>,
}
unsafe impl Send for B200 {}
unsafe impl Sync for B200 {}

impl B200 {
    /// There is no CPU fallback: without a CUDA device this fails (`pb2_ctx_create` -> `PB2_ERR_CUDA`).
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut ctx = core::ptr::null_mut();
        let st = unsafe { sys::pb2_ctx_create(device, &mut ctx) };
        if st != sys::PB2_OK {
            return Err(Error::Cuda(format!("pb2_ctx_create({device}) failed with status {st}")));
        }
        Ok(B200 { ctx: Mutex::new(ctx) })
    }

    /// Runs `f` with the raw context locked and maps its status.
    fn call(&self, f: impl FnOnce(*mut sys::pb2_ctx) -> i32) -> Result<(), Error> {
        let guard = self.ctx.lock().unwrap();
        let st = f(*guard);
        Self::map_status(*guard, st, 0)
    }

    fn map_status(ctx: *mut sys::pb2_ctx, st: i32, required: u64) -> Result<(), Error> {
        if st == sys::PB2_OK {
            return Ok(());
        }
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::pb2_last_error(ctx)) }.to_string_lossy().into_owned();
        Err(match st {
            sys::PB2_ERR_INVALID => Error::Invalid(msg),
            sys::PB2_ERR_OVERFLOW => Error::Overflow(required),
            sys::PB2_ERR_UNSUPPORTED => Error::Unsupported,
            sys::PB2_ERR_DEPTH => Error::Depth(msg),
            _ => Error::Cuda(msg),
        })
    }

    /// Calls with a caller-owned output list (`cap` elements in, `count` out): grows the list and calls again on
    /// `PB2_ERR_OVERFLOW`, which reports the required size in `count` and leaves the first `cap` entries valid.
    fn call_growing<T: Clone + Default>(
        &self,
        out: &mut Vec<T>,
        mut f: impl FnMut(*mut sys::pb2_ctx, *mut T, u64, &mut u64) -> i32,
    ) -> Result<(), Error> {
        let guard = self.ctx.lock().unwrap();
        loop {
            let mut count = 0u64;
            let st = f(*guard, out.as_mut_ptr(), out.len() as u64, &mut count);
            if st == sys::PB2_ERR_OVERFLOW {
                out.resize(count as usize, T::default());
                continue;
            }
            Self::map_status(*guard, st, count)?;
            out.truncate(count as usize);
            return Ok(());
        }
    }

    pub fn synchronize(&self) -> Result<(), Error> {
        self.call(|c| unsafe { sys::pb2_ctx_synchronize(c) })
    }
}

impl Drop for B200 {
    fn drop(&mut self) {
        let ctx = *self.ctx.lock().unwrap();
        if !ctx.is_null() {
            unsafe { sys::pb2_ctx_destroy(ctx) };
        }
    }
}

#[inline]
fn iso7(m: &Isometry<Real>) -> [f32; 7] {
    // Isometry3 = { UnitQuaternion [i, j, k, w], Translation3 [x, y, z] } (nalgebra field order: rotation, translation)
    let q = m.rotation.coords;
    let t = m.translation.vector;
    [q.x, q.y, q.z, q.w, t.x, t.y, t.z]
}
const IDENTITY7: [f32; 7] = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0];

// ------------------------------------------------------------------------------------------------ partitioning::Bvh
/// `parry3d::partitioning::Bvh` on the device. Leaf `i` is the i-th box handed to `from_leaves` (bvh_tree.rs:1835).
pub struct Bvh<'c> {
    c: &'c B200,
    h: *mut sys::pb2_bvh,
}

impl<'c> Bvh<'c> {
    /// `Bvh::from_leaves(strategy, leaves)` — bvh_tree.rs:1835. `Binned` links a Morton LBVH, `Ploc` runs the reference's own
    /// PLOC rule on the GPU; query results do not depend on the strategy.
    pub fn from_leaves(c: &'c B200, strategy: BvhBuildStrategy, leaves: &[Aabb]) -> Result<Self, Error> {
        let mut h = core::ptr::null_mut();
        let s = match strategy {
            BvhBuildStrategy::Binned => sys::PB2_BUILD_BINNED,
            BvhBuildStrategy::Ploc => sys::PB2_BUILD_PLOC,
        };
        // Aabb is #[repr(C)] { mins: Point3<f32>, maxs: Point3<f32> } = 6 contiguous f32 (aabb.rs:110)
        c.call(|ctx| unsafe { sys::pb2_bvh_build(ctx, leaves.as_ptr() as *const f32, leaves.len() as u32, s, sys::PB2_MEM_HOST, &mut h) })?;
        Ok(Bvh { c, h })
    }

    /// `Bvh::from_iter(strategy, iter of (leaf index, Aabb))` — bvh_tree.rs:1891: indices may have gaps; missing slots are
    /// inert (`Aabb::new_invalid`) leaves.
    pub fn from_iter(c: &'c B200, strategy: BvhBuildStrategy, leaves: impl IntoIterator<Item = (usize, Aabb)>) -> Result<Self, Error> {
        let items: Vec<(usize, Aabb)> = leaves.into_iter().collect();
        let n = items.iter().map(|(i, _)| i + 1).max().unwrap_or(0);
        let mut boxes = vec![Aabb::new_invalid(); n];
        for (i, a) in items {
            boxes[i] = a;
        }
        Self::from_leaves(c, strategy, &boxes)
    }

    pub fn leaf_count(&self) -> u32 {
        unsafe { sys::pb2_bvh_leaf_count(self.h) }
    }

    /// `Bvh::insert_or_update_partially(aabb, leaf_index, change_detection_margin)` — bvh_insert.rs:209-231, for all the leaves
    /// that moved this frame in one call. With a margin > 0 a leaf whose stored box still contains the new one is left alone,
    /// otherwise it is enlarged by the margin and flagged as changed.
    pub fn insert_or_update_partially(&mut self, aabbs: &[Aabb], leaf_indices: &[u32], change_detection_margin: Real) -> Result<(), Error> {
        assert_eq!(aabbs.len(), leaf_indices.len());
        let h = self.h;
        self.c.call(|ctx| unsafe {
            sys::pb2_bvh_update_leaves(ctx, h, leaf_indices.as_ptr(), aabbs.as_ptr() as *const f32, aabbs.len() as u32, change_detection_margin, sys::PB2_MEM_HOST)
        })
    }

    /// `Bvh::refit(&mut workspace)` — bvh_refit.rs:170 (no workspace needed: scratch lives in the context).
    pub fn refit(&mut self) -> Result<(), Error> {
        let h = self.h;
        self.c.call(|ctx| unsafe { sys::pb2_bvh_refit(ctx, h) })
    }

    /// `Bvh::refit_without_opt` — bvh_refit.rs:326: same boxes and flags; the DFS re-layout the two differ by is a CPU cache matter.
    pub fn refit_without_opt(&mut self) -> Result<(), Error> {
        self.refit()
    }

    /// `Bvh::rebuild(&mut workspace, strategy)` — bvh_binned_build.rs:11-36: same leaves with their change flags, new topology.
    pub fn rebuild(&mut self, strategy: BvhBuildStrategy) -> Result<(), Error> {
        let h = self.h;
        let s = if matches!(strategy, BvhBuildStrategy::Ploc) { sys::PB2_BUILD_PLOC } else { sys::PB2_BUILD_BINNED };
        self.c.call(|ctx| unsafe { sys::pb2_bvh_rebuild(ctx, h, s) })
    }

    /// `Bvh::optimize_incremental` — bvh_optimize.rs:237: the reference re-bins ~5 % of the leaves per frame to keep a refitted
    /// tree from degrading; a full rebuild costs ~1 ms per million leaves here, so the whole tree is rebuilt.
    pub fn optimize_incremental(&mut self) -> Result<(), Error> {
        self.rebuild(BvhBuildStrategy::Binned)
    }

    /// `Bvh::insert(aabb, leaf_index)` — bvh_insert.rs:126, batched: new indices grow the leaf-id space, then update + rebuild
    /// (structural edits are whole-tree rebuilds on the GPU instead of the SAH descent with rotations).
    pub fn insert(&mut self, aabbs: &[Aabb], leaf_indices: &[u32]) -> Result<(), Error> {
        let need = leaf_indices.iter().copied().max().map(|m| m + 1).unwrap_or(0);
        let h = self.h;
        if need > self.leaf_count() {
            self.c.call(|ctx| unsafe { sys::pb2_bvh_resize(ctx, h, need) })?;
        }
        self.insert_or_update_partially(aabbs, leaf_indices, 0.0)?;
        self.rebuild(BvhBuildStrategy::Binned)
    }

    /// `Bvh::remove(leaf_index)` — bvh_tree.rs:2360, batched; unknown indices are ignored like in the reference.
    pub fn remove(&mut self, leaf_indices: &[u32]) -> Result<(), Error> {
        let h = self.h;
        self.c.call(|ctx| unsafe { sys::pb2_bvh_remove_leaves(ctx, h, leaf_indices.as_ptr(), leaf_indices.len() as u32, sys::PB2_MEM_HOST) })
    }

    /// `Bvh::root_aabb` — bvh_tree.rs:1991.
    pub fn root_aabb(&self) -> Result<Aabb, Error> {
        let mut a = [0f32; 6];
        let h = self.h;
        self.c.call(|ctx| unsafe { sys::pb2_bvh_root_aabb(ctx, h, a.as_mut_ptr()) })?;
        Ok(Aabb::new(Point::new(a[0], a[1], a[2]), Point::new(a[3], a[4], a[5])))
    }

    /// `Bvh::intersect_aabb(&aabb)` — bvh_queries.rs:203, for a batch of query boxes: CSR (offsets, leaf ids). The order inside a
    /// query's group is unspecified, like the reference's iterator order depends on its tree.
    pub fn intersect_aabbs(&self, queries: &[Aabb]) -> Result<(Vec<u32>, Vec<u32>), Error> {
        let mut offsets = vec![0u32; queries.len() + 1];
        let mut ids = vec![0u32; 8 * queries.len() + 64];
        let h = self.h;
        self.c.call_growing(&mut ids, |ctx, out, cap, count| unsafe {
            sys::pb2_bvh_intersect_aabbs(ctx, h, queries.as_ptr() as *const f32, queries.len() as u32, offsets.as_mut_ptr(), out, cap, count, sys::PB2_MEM_HOST)
        })?;
        Ok((offsets, ids))
    }

    /// Single-query form with the reference's signature.
    pub fn intersect_aabb(&self, aabb: &Aabb) -> Result<impl Iterator<Item = u32>, Error> {
        let (_, ids) = self.intersect_aabbs(core::slice::from_ref(aabb))?;
        Ok(ids.into_iter())
    }

    /// `Bvh::traverse_bvtt_single_tree::<CHANGE_DETECTION>(&mut workspace, &mut f)` — bvh_traverse_bvtt.rs:19: `f(a, b)` once per
    /// unordered pair of leaves with overlapping boxes (with change detection: pairs touching a leaf flagged by the last refit).
    pub fn traverse_bvtt_single_tree<const CHANGE_DETECTION: bool>(&self, f: &mut impl FnMut(u32, u32)) -> Result<(), Error> {
        let mut pairs = vec![[0u32; 2]; 8 * self.leaf_count() as usize + 1024];
        let h = self.h;
        self.c.call_growing(&mut pairs, |ctx, out, cap, count| unsafe {
            sys::pb2_bvh_self_pairs(ctx, h, CHANGE_DETECTION as i32, out as *mut u32, cap, count, sys::PB2_MEM_HOST)
        })?;
        for p in &pairs {
            f(p[0], p[1]);
        }
        Ok(())
    }

    /// `Bvh::leaf_pairs(&other, |a, b| a.intersects(b))` — bvh_traverse_bvtt.rs:210 (the check is fixed to box overlap: a closure
    /// cannot run on the device).
    pub fn leaf_pairs(&self, other: &Bvh<'c>) -> Result<Vec<[u32; 2]>, Error> {
        let mut pairs = vec![[0u32; 2]; 8 * self.leaf_count().max(other.leaf_count()) as usize + 1024];
        let (a, b) = (self.h, other.h);
        self.c.call_growing(&mut pairs, |ctx, out, cap, count| unsafe { sys::pb2_bvh_leaf_pairs(ctx, a, b, out as *mut u32, cap, count, sys::PB2_MEM_HOST) })?;
        Ok(pairs)
    }

    /// The node array in the reference's own layout (`BvhNodeWide`, 64 bytes, bvh_tree.rs:263-266) + `parents` +
    /// `leaf_node_indices`, so that the tree stays inspectable / serialisable with parry's types.
    pub fn download(&self) -> Result<(Vec<BvhNodeWide>, Vec<u32>, Vec<u32>), Error> {
        let n_nodes = unsafe { sys::pb2_bvh_node_count(self.h) } as usize;
        let mut nodes: Vec<BvhNodeWide> = Vec::with_capacity(n_nodes);
        let mut parents = vec![0u32; n_nodes];
        let mut leaf_node_indices = vec![0u32; self.leaf_count() as usize];
        let h = self.h;
        self.c.call(|ctx| unsafe {
            sys::pb2_bvh_download(ctx, h, nodes.as_mut_ptr() as *mut c_void, parents.as_mut_ptr(), leaf_node_indices.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        unsafe { nodes.set_len(n_nodes) };
        Ok((nodes, parents, leaf_node_indices))
    }

    /// `Bvh::cast_ray(&ray, max_toi, |leaf, best| shapes[leaf].cast_ray(..))` — bvh_queries.rs:260 with the leaf callback replaced
    /// by typed leaves: leaf `i` is `table` shape `shape_ids[i]` at `poses[i]`. `None` = miss.
    pub fn cast_rays(&self, table: &ShapeTable<'c>, shape_ids: &[u32], poses: &[Isometry<Real>], rays: &[Ray], max_time_of_impact: Real, solid: bool)
                     -> Result<Vec<Option<(u32, RayIntersection)>>, Error> {
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let m = rays.len();
        let (mut toi, mut leaf, mut normal, mut feat) = (vec![0f32; m], vec![0u32; m], vec![[0f32; 3]; m], vec![0u32; m]);
        let (h, t) = (self.h, table.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_bvh_cast_rays_shapes(ctx, h, t, shape_ids.as_ptr(), p7.as_ptr() as *const f32, rays.as_ptr() as *const f32, m as u32, max_time_of_impact,
                                          solid as i32, toi.as_mut_ptr(), leaf.as_mut_ptr(), normal.as_mut_ptr() as *mut f32, feat.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        Ok((0..m).map(|i| (leaf[i] != sys::PB2_INVALID_U32).then(|| {
            let feature = if feat[i] == sys::PB2_FEATURE_UNKNOWN { FeatureId::Unknown } else { FeatureId::Face(feat[i]) };
            (leaf[i], RayIntersection::new(toi[i], Vector::new(normal[i][0], normal[i][1], normal[i][2]), feature))
        })).collect())
    }
}

impl<'c> Bvh<'c> {
    /// `Bvh::project_point(point, max_distance, |leaf, _| shape(leaf).project_point(pose(leaf), point, solid))`
    /// (bvh_queries.rs:213-227) for a batch of points over typed leaves: `Some((leaf, (distance, projection)))` per point.
    /// `Err(Unsupported)`: `solid == false` and the point lies inside a ConvexPolyhedron leaf (the reference's EPA branch).
    pub fn project_points(&self, table: &ShapeTable<'c>, shape_ids: &[u32], poses: &[Isometry<Real>], points: &[Point<Real>], max_distance: Real,
                          solid: bool) -> Result<Vec<Result<Option<(u32, (Real, parry3d::query::PointProjection))>, Unsupported>>, Error> {
        let m = points.len();
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let q: Vec<[f32; 3]> = points.iter().map(|p| [p.x, p.y, p.z]).collect();
        let (mut proj, mut inside, mut leaf, mut status) = (vec![[0f32; 3]; m], vec![0u8; m], vec![0u32; m], vec![0u8; m]);
        let (h, t) = (self.h, table.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_bvh_project_points_shapes(ctx, h, t, shape_ids.as_ptr(), p7.as_ptr() as *const f32, q.as_ptr() as *const f32, m as u32, max_distance,
                                               solid as i32, proj.as_mut_ptr() as *mut f32, inside.as_mut_ptr(), leaf.as_mut_ptr(), status.as_mut_ptr(),
                                               sys::PB2_MEM_HOST)
        })?;
        Ok((0..m).map(|k| match status[k] {
            0 => Ok(None),
            1 => {
                let pt = Point::new(proj[k][0], proj[k][1], proj[k][2]);
                Ok(Some((leaf[k], (nalgebra::distance(&pt, &points[k]), parry3d::query::PointProjection::new(inside[k] != 0, pt)))))
            }
            _ => Err(Unsupported),
        }).collect())
    }
}

impl Drop for Bvh<'_> {
    fn drop(&mut self) {
        let h = self.h;
        let _ = self.c.call(|ctx| unsafe { sys::pb2_bvh_destroy(ctx, h) });
    }
}

// ------------------------------------------------------------------------------------------------ shape::TriMesh + RayCast
/// `parry3d::shape::TriMesh` (vertices + indices + its Bvh, trimesh.rs:519-529) on the device.
pub struct TriMesh<'c> {
    c: &'c B200,
    h: *mut sys::pb2_trimesh,
    num_triangles: u32,
}

/// `TriMesh::cast_ray_with_culling`'s mode (ray_trimesh.rs:50-56).
#[derive(Copy, Clone, PartialEq, Eq, Debug)]
pub enum RayCullingMode {
    IgnoreBackfaces,
    IgnoreFrontfaces,
}

impl<'c> TriMesh<'c> {
    /// `TriMesh::new(vertices, indices)` — trimesh.rs:607. Empty index buffers and out-of-range indices are errors
    /// (`TriMeshBuilderError::EmptyIndices`).
    pub fn new(c: &'c B200, vertices: &[Point<Real>], indices: &[[u32; 3]]) -> Result<Self, Error> {
        let mut h = core::ptr::null_mut();
        c.call(|ctx| unsafe {
            sys::pb2_trimesh_create(ctx, vertices.as_ptr() as *const f32, vertices.len() as u32, indices.as_ptr() as *const u32, indices.len() as u32, sys::PB2_MEM_HOST, &mut h)
        })?;
        Ok(TriMesh { c, h, num_triangles: indices.len() as u32 })
    }

    fn cast(&self, m: Option<&Isometry<Real>>, rays: &[Ray], max_toi: Real, culling: Option<RayCullingMode>, want_normal: bool)
            -> Result<(Vec<f32>, Vec<u32>, Vec<[f32; 3]>, Vec<u32>), Error> {
        let n = rays.len();
        let pose = m.map(iso7);
        let pose_ptr = pose.as_ref().map_or(core::ptr::null(), |p| p.as_ptr());
        let (mut toi, mut tri) = (vec![0f32; n], vec![0u32; n]);
        let (mut normal, mut feat) = if want_normal { (vec![[0f32; 3]; n], vec![0u32; n]) } else { (Vec::new(), Vec::new()) };
        let (np, fp) = if want_normal { (normal.as_mut_ptr() as *mut f32, feat.as_mut_ptr()) } else { (core::ptr::null_mut(), core::ptr::null_mut()) };
        let h = self.h;
        // Ray is #[repr(C)] { origin: Point3<f32>, dir: Vector3<f32> } = 6 contiguous f32 (ray.rs:74-88)
        self.c.call(|ctx| unsafe {
            match culling {
                None => sys::pb2_trimesh_cast_rays(ctx, h, pose_ptr, rays.as_ptr() as *const f32, n as u32, max_toi, 1, toi.as_mut_ptr(), tri.as_mut_ptr(), np, fp, sys::PB2_MEM_HOST),
                Some(mode) => {
                    let k = if mode == RayCullingMode::IgnoreBackfaces { sys::PB2_CULL_IGNORE_BACKFACES } else { sys::PB2_CULL_IGNORE_FRONTFACES } as i32;
                    sys::pb2_trimesh_cast_rays_with_culling(ctx, h, pose_ptr, rays.as_ptr() as *const f32, n as u32, max_toi, k, toi.as_mut_ptr(), tri.as_mut_ptr(), np, fp, sys::PB2_MEM_HOST)
                }
            }
        })?;
        Ok((toi, tri, normal, feat))
    }

    /// `RayCast::cast_ray(m, ray, max_time_of_impact, solid)` — ray.rs:381 (`cast_local_ray` with `m = None`), one entry per ray.
    /// `solid` is ignored by the 3D triangle test, as in the reference (ray_triangle.rs:53).
    pub fn cast_rays(&self, m: Option<&Isometry<Real>>, rays: &[Ray], max_time_of_impact: Real, _solid: bool) -> Result<Vec<Option<Real>>, Error> {
        let (toi, tri, _, _) = self.cast(m, rays, max_time_of_impact, None, false)?;
        Ok(toi.iter().zip(&tri).map(|(t, k)| (*k != sys::PB2_INVALID_U32).then_some(*t)).collect())
    }

    /// `RayCast::cast_ray_and_get_normal` — ray.rs:393: feature = `Face(i)` for a front-face hit of triangle i, `Face(i + num_triangles)`
    /// for a back-face hit (ray_trimesh.rs:28-32).
    pub fn cast_rays_and_get_normal(&self, m: Option<&Isometry<Real>>, rays: &[Ray], max_time_of_impact: Real, _solid: bool)
                                    -> Result<Vec<Option<RayIntersection>>, Error> {
        let (toi, tri, normal, feat) = self.cast(m, rays, max_time_of_impact, None, true)?;
        Ok((0..rays.len()).map(|i| (tri[i] != sys::PB2_INVALID_U32)
            .then(|| RayIntersection::new(toi[i], Vector::new(normal[i][0], normal[i][1], normal[i][2]), FeatureId::Face(feat[i])))).collect())
    }

    /// `TriMesh::cast_ray_with_culling` — ray_trimesh.rs:139.
    pub fn cast_rays_with_culling(&self, m: Option<&Isometry<Real>>, rays: &[Ray], max_time_of_impact: Real, culling: RayCullingMode)
                                  -> Result<Vec<Option<RayIntersection>>, Error> {
        let (toi, tri, normal, feat) = self.cast(m, rays, max_time_of_impact, Some(culling), true)?;
        Ok((0..rays.len()).map(|i| (tri[i] != sys::PB2_INVALID_U32)
            .then(|| RayIntersection::new(toi[i], Vector::new(normal[i][0], normal[i][1], normal[i][2]), FeatureId::Face(feat[i])))).collect())
    }

    /// Index of the hit triangle next to each toi (`CompositeShapeRef::cast_local_ray` returns it, ray_composite_shape.rs:20).
    pub fn cast_rays_with_ids(&self, m: Option<&Isometry<Real>>, rays: &[Ray], max_time_of_impact: Real) -> Result<Vec<Option<(u32, Real)>>, Error> {
        let (toi, tri, _, _) = self.cast(m, rays, max_time_of_impact, None, false)?;
        Ok(toi.iter().zip(&tri).map(|(t, k)| (*k != sys::PB2_INVALID_U32).then_some((*k, *t))).collect())
    }

    pub fn num_triangles(&self) -> u32 {
        self.num_triangles
    }

    /// `query::contact(mesh_pose, &trimesh, poses[k], shape k, prediction)` — the composite-shape arm
    /// (contact_composite_shape_shape.rs:14-61): `(contact, triangle)` per collider.
    pub fn contact_shapes(&self, mesh_pose: &Isometry<Real>, table: &ShapeTable<'c>, shape_ids: &[u32], poses: &[Isometry<Real>], prediction: Real)
                          -> Result<Vec<Result<Option<(Contact, u32)>, Unsupported>>, Error> {
        let n = shape_ids.len();
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let mp = iso7(mesh_pose);
        let (mut out, mut status, mut part) = (vec![sys::pb2_contact::default(); n], vec![0u8; n], vec![0u32; n]);
        let (h, t) = (self.h, table.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_trimesh_contact_shapes(ctx, h, mp.as_ptr(), t, shape_ids.as_ptr(), p7.as_ptr() as *const f32, n as u32, prediction, out.as_mut_ptr(), status.as_mut_ptr(),
                                            part.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| contact_of_status(status[k], &out[k]).map(|c| c.map(|c| (c, part[k])))).collect())
    }
}

impl<'c> TriMesh<'c> {
    /// `query::cast_shapes(mesh_pose, mesh_vel, &trimesh, poses[k], vels[k], shape k, options)` — the composite arm
    /// (shape_cast_composite_shape_shape.rs:65-83) — or, with `mesh_second`, the shape as shape 1 and the swapped hit (:86-105):
    /// `(hit, triangle)` per collider. `options.stop_at_penetration` must be true (the reference's default).
    pub fn cast_shapes(&self, mesh_pose: &Isometry<Real>, mesh_vel: &Vector<Real>, table: &ShapeTable<'c>, shape_ids: &[u32], poses: &[Isometry<Real>],
                       vels: &[Vector<Real>], mesh_second: bool, options: ShapeCastOptions)
                       -> Result<Vec<Result<Option<(ShapeCastHit, u32)>, Unsupported>>, Error> {
        let n = shape_ids.len();
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let v3: Vec<[f32; 3]> = vels.iter().map(|v| [v.x, v.y, v.z]).collect();
        let (mp, mv) = (iso7(mesh_pose), [mesh_vel.x, mesh_vel.y, mesh_vel.z]);
        let (mut out, mut status, mut part) = (vec![0f32; 13 * n], vec![0u8; n], vec![0u32; n]);
        let (h, t) = (self.h, table.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_trimesh_cast_shapes(ctx, h, mp.as_ptr(), mv.as_ptr(), t, shape_ids.as_ptr(), p7.as_ptr() as *const f32, v3.as_ptr() as *const f32,
                                         mesh_second as i32, options.max_time_of_impact, options.target_distance, options.stop_at_penetration as i32,
                                         options.compute_impact_geometry_on_penetration as i32, n as u32, out.as_mut_ptr(), status.as_mut_ptr(),
                                         part.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| cast_of_status(status[k], &out[13 * k..13 * k + 13]).map(|h| h.map(|h| (h, part[k])))).collect())
    }

    /// `query::distance(mesh_pose, &trimesh, poses[k], shape k)` — or with `mesh_second` the shape first — for every k: the composite
    /// arms of `DefaultQueryDispatcher::distance` (distance_composite_shape_shape.rs:46-77). `Err(Unsupported)`: unknown shape id.
    pub fn distance_shapes(&self, mesh_pose: &Isometry<Real>, table: &ShapeTable<'c>, shape_ids: &[u32], poses: &[Isometry<Real>], mesh_second: bool)
                           -> Result<Vec<Result<Real, Unsupported>>, Error> {
        let n = shape_ids.len();
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let mp = iso7(mesh_pose);
        let (mut dist, mut status, mut part) = (vec![0f32; n], vec![0u8; n], vec![0u32; n]);
        let (h, t) = (self.h, table.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_trimesh_distance_shapes(ctx, h, mp.as_ptr(), t, shape_ids.as_ptr(), p7.as_ptr() as *const f32, mesh_second as i32, n as u32,
                                             dist.as_mut_ptr(), status.as_mut_ptr(), part.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| if status[k] == 0 { Ok(dist[k]) } else { Err(Unsupported) }).collect())
    }

    /// `query::cast_shapes(pos1[k], vel1[k], &self, pos2[k], vel2[k], &other, options)` for two TriMeshes (the nesting of the
    /// reference's tests/geometry/trimesh_trimesh_toi.rs): `(hit, [triangle of self, triangle of other])`.
    pub fn cast_trimesh(&self, pos1: &[Isometry<Real>], vel1: &[Vector<Real>], other: &TriMesh<'c>, pos2: &[Isometry<Real>], vel2: &[Vector<Real>],
                        options: ShapeCastOptions) -> Result<Vec<Result<Option<(ShapeCastHit, [u32; 2])>, Unsupported>>, Error> {
        let n = pos1.len();
        let (p1, p2): (Vec<[f32; 7]>, Vec<[f32; 7]>) = (pos1.iter().map(iso7).collect(), pos2.iter().map(iso7).collect());
        let v1: Vec<[f32; 3]> = vel1.iter().map(|v| [v.x, v.y, v.z]).collect();
        let v2: Vec<[f32; 3]> = vel2.iter().map(|v| [v.x, v.y, v.z]).collect();
        let (mut out, mut status, mut parts) = (vec![0f32; 13 * n], vec![0u8; n], vec![[0u32; 2]; n]);
        let (h1, h2) = (self.h, other.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_trimesh_cast_trimesh(ctx, h1, p1.as_ptr() as *const f32, v1.as_ptr() as *const f32, h2, p2.as_ptr() as *const f32,
                                          v2.as_ptr() as *const f32, options.max_time_of_impact, options.target_distance,
                                          options.stop_at_penetration as i32, options.compute_impact_geometry_on_penetration as i32, n as u32,
                                          out.as_mut_ptr(), status.as_mut_ptr(), parts.as_mut_ptr() as *mut u32, sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| cast_of_status(status[k], &out[13 * k..13 * k + 13]).map(|h| h.map(|h| (h, parts[k])))).collect())
    }
}

impl Drop for TriMesh<'_> {
    fn drop(&mut self) {
        let h = self.h;
        let _ = self.c.call(|ctx| unsafe { sys::pb2_trimesh_destroy(ctx, h) });
    }
}

// ------------------------------------------------------------------------------------------------ shapes
/// Typed shape table (`pb2_shapes`): the Ball / Cuboid / ConvexPolyhedron instances the device kernels know. Shapes are
/// identified by the address of the `dyn Shape` they were registered from, which is all a `QueryDispatcher` call gets to see.
pub struct ShapeTable<'c> {
    c: &'c B200,
    h: *mut sys::pb2_shapes,
    ids: HashMap<usize, u32>,
}

fn shape_key(s: &dyn Shape) -> usize {
    s as *const dyn Shape as *const () as usize
}

impl<'c> ShapeTable<'c> {
    /// Registers every shape of `shapes` that the path supports; the others are simply absent from the table (queries on them
    /// answer `Unsupported` and fall through the dispatcher chain). The shapes must outlive the table and not move.
    pub fn new(c: &'c B200, shapes: &[&dyn Shape]) -> Result<Self, Error> {
        let (mut kinds, mut params, mut points, mut ids) = (Vec::<u8>::new(), Vec::<[f32; 4]>::new(), Vec::<[f32; 3]>::new(), HashMap::new());
        for s in shapes {
            let (kind, p) = match s.as_typed_shape() {
                TypedShape::Ball(b) => (sys::PB2_SHAPE_BALL as u8, [b.radius, 0.0, 0.0, 0.0]),
                TypedShape::Cuboid(cb) => (sys::PB2_SHAPE_CUBOID as u8, [cb.half_extents.x, cb.half_extents.y, cb.half_extents.z, 0.0]),
                TypedShape::ConvexPolyhedron(cp) => {
                    let first = points.len() as u32;
                    points.extend(cp.points().iter().map(|p| [p.x, p.y, p.z]));
                    // params of a hull = {first point, point count} as raw u32 bits (include/parry_b200.h: pb2_shapes_create)
                    (sys::PB2_SHAPE_CONVEX as u8, [f32::from_bits(first), f32::from_bits(cp.points().len() as u32), 0.0, 0.0])
                }
                _ => continue,
            };
            ids.insert(shape_key(*s), kinds.len() as u32);
            kinds.push(kind);
            params.push(p);
        }
        let mut h = core::ptr::null_mut();
        c.call(|ctx| unsafe {
            sys::pb2_shapes_create(ctx, kinds.as_ptr(), params.as_ptr() as *const f32, kinds.len() as u32, points.as_ptr() as *const f32, points.len() as u32, &mut h)
        })?;
        Ok(ShapeTable { c, h, ids })
    }

    pub fn id_of(&self, s: &dyn Shape) -> Option<u32> {
        self.ids.get(&shape_key(s)).copied()
    }

    /// `Shape::compute_aabb(position)` for a batch (aabb_ball.rs:8, aabb_cuboid.rs:9, aabb_convex_polyhedron.rs:8).
    pub fn compute_aabbs(&self, shape_ids: &[u32], poses: &[Isometry<Real>]) -> Result<Vec<Aabb>, Error> {
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let mut out = vec![Aabb::new_invalid(); shape_ids.len()];
        let h = self.h;
        self.c.call(|ctx| unsafe {
            sys::pb2_shapes_compute_aabbs(ctx, h, shape_ids.as_ptr(), p7.as_ptr() as *const f32, shape_ids.len() as u32, out.as_mut_ptr() as *mut f32, sys::PB2_MEM_HOST)
        })?;
        Ok(out)
    }
}

impl Drop for ShapeTable<'_> {
    fn drop(&mut self) {
        let h = self.h;
        let _ = self.c.call(|ctx| unsafe { sys::pb2_shapes_destroy(ctx, h) });
    }
}

// ------------------------------------------------------------------------------------------------ shape::Compound
/// A table of `Compound`s (`pb2_compounds`, shape/compound.rs:113-144) whose parts are shapes of one [`ShapeTable`]. The composite
/// arms of `DefaultQueryDispatcher::contact` (default_query_dispatcher.rs:338-351 -> contact_composite_shape_shape.rs:14-76) in
/// batches; every method returns, next to the contact, the winning part(s) the reference's traversal would have reported.
pub struct Compounds<'c> {
    c: &'c B200,
    h: *mut sys::pb2_compounds,
}
unsafe impl Send for Compounds<'_> {}
unsafe impl Sync for Compounds<'_> {}

impl<'c> Compounds<'c> {
    /// Every part must be a shape `table` holds (looked up by address, like `ShapeTable::id_of`); otherwise `Error::Unsupported`.
    pub fn new(c: &'c B200, table: &ShapeTable<'c>, compounds: &[&parry3d::shape::Compound]) -> Result<Self, Error> {
        let (mut first, mut count, mut part_shape, mut part_pose) = (Vec::<u32>::new(), Vec::<u32>::new(), Vec::<u32>::new(), Vec::<[f32; 7]>::new());
        for cp in compounds {
            first.push(part_shape.len() as u32);
            count.push(cp.shapes().len() as u32);
            for (pose, shape) in cp.shapes() {
                part_shape.push(table.id_of(shape.as_ref()).ok_or(Error::Unsupported)?);
                part_pose.push(iso7(pose));
            }
        }
        let mut h = core::ptr::null_mut();
        let t = table.h;
        c.call(|ctx| unsafe {
            sys::pb2_compounds_create(ctx, t, first.as_ptr(), count.as_ptr(), first.len() as u32, part_shape.as_ptr(), part_pose.as_ptr() as *const f32,
                                      part_shape.len() as u32, &mut h)
        })?;
        Ok(Compounds { c, h })
    }

    /// `query::contact(compound_poses[k], compound k, shape_poses[k], shape k, prediction)`, or with `compound_second` the shape
    /// first (contact_shape_composite_shape, :63-76): `(contact, winning part)`.
    pub fn contact_shapes(&self, compound_ids: &[u32], compound_poses: &[Isometry<Real>], shape_ids: &[u32], shape_poses: &[Isometry<Real>],
                          prediction: Real, compound_second: bool) -> Result<Vec<Result<Option<(Contact, u32)>, Unsupported>>, Error> {
        let n = compound_ids.len();
        let (pc, ps): (Vec<[f32; 7]>, Vec<[f32; 7]>) = (compound_poses.iter().map(iso7).collect(), shape_poses.iter().map(iso7).collect());
        let (mut out, mut status, mut part) = (vec![sys::pb2_contact::default(); n], vec![0u8; n], vec![0u32; n]);
        let h = self.h;
        self.c.call(|ctx| unsafe {
            sys::pb2_compound_contact_shapes(ctx, h, compound_ids.as_ptr(), pc.as_ptr() as *const f32, shape_ids.as_ptr(), ps.as_ptr() as *const f32,
                                             n as u32, prediction, compound_second as i32, out.as_mut_ptr(), status.as_mut_ptr(), part.as_mut_ptr(),
                                             sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| contact_of_status(status[k], &out[k]).map(|c| c.map(|c| (c, part[k])))).collect())
    }

    /// `query::contact(poses1[k], compound ids1[k], poses2[k], compound ids2[k], prediction)`: `(contact, [part of 1, part of 2])`.
    pub fn contact_compounds(&self, ids1: &[u32], poses1: &[Isometry<Real>], ids2: &[u32], poses2: &[Isometry<Real>], prediction: Real)
                             -> Result<Vec<Result<Option<(Contact, [u32; 2])>, Unsupported>>, Error> {
        let n = ids1.len();
        let (p1, p2): (Vec<[f32; 7]>, Vec<[f32; 7]>) = (poses1.iter().map(iso7).collect(), poses2.iter().map(iso7).collect());
        let (mut out, mut status, mut parts) = (vec![sys::pb2_contact::default(); n], vec![0u8; n], vec![[0u32; 2]; n]);
        let h = self.h;
        self.c.call(|ctx| unsafe {
            sys::pb2_compound_contact_compounds(ctx, h, ids1.as_ptr(), p1.as_ptr() as *const f32, ids2.as_ptr(), p2.as_ptr() as *const f32, n as u32,
                                                prediction, out.as_mut_ptr(), status.as_mut_ptr(), parts.as_mut_ptr() as *mut u32, sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| contact_of_status(status[k], &out[k]).map(|c| c.map(|c| (c, parts[k])))).collect())
    }

    /// `query::contact(poses[k], compound ids[k], mesh_pose, &trimesh, prediction)` or, with `mesh_first`,
    /// `query::contact(mesh_pose, &trimesh, poses[k], compound ids[k], prediction)`: `(contact, [part, triangle])`.
    pub fn contact_trimesh(&self, ids: &[u32], poses: &[Isometry<Real>], mesh: &TriMesh<'c>, mesh_pose: &Isometry<Real>, prediction: Real, mesh_first: bool)
                           -> Result<Vec<Result<Option<(Contact, [u32; 2])>, Unsupported>>, Error> {
        let n = ids.len();
        let p7: Vec<[f32; 7]> = poses.iter().map(iso7).collect();
        let mp = iso7(mesh_pose);
        let (mut out, mut status, mut parts) = (vec![sys::pb2_contact::default(); n], vec![0u8; n], vec![[0u32; 2]; n]);
        let (h, m) = (self.h, mesh.h);
        self.c.call(|ctx| unsafe {
            sys::pb2_compound_contact_trimesh(ctx, h, ids.as_ptr(), p7.as_ptr() as *const f32, m, mp.as_ptr(), n as u32, prediction, mesh_first as i32,
                                              out.as_mut_ptr(), status.as_mut_ptr(), parts.as_mut_ptr() as *mut u32, sys::PB2_MEM_HOST)
        })?;
        Ok((0..n).map(|k| contact_of_status(status[k], &out[k]).map(|c| c.map(|c| (c, parts[k])))).collect())
    }
}

impl Drop for Compounds<'_> {
    fn drop(&mut self) {
        let h = self.h;
        let _ = self.c.call(|ctx| unsafe { sys::pb2_compounds_destroy(ctx, h) });
    }
}

/// `PB2_CAST_*` + the 13 floats of one hit -> `cast_shapes`' return value. Unknown shape, or a documented host case: `Unsupported`,
/// so that a chain re-runs the pair on the CPU.
fn cast_of_status(status: u8, out: &[f32]) -> Result<Option<ShapeCastHit>, Unsupported> {
    let hit = |st: ShapeCastStatus| ShapeCastHit {
        witness1: Point::new(out[0], out[1], out[2]),
        witness2: Point::new(out[3], out[4], out[5]),
        normal1: nalgebra::Unit::new_unchecked(Vector::new(out[6], out[7], out[8])),
        normal2: nalgebra::Unit::new_unchecked(Vector::new(out[9], out[10], out[11])),
        time_of_impact: out[12],
        status: st,
    };
    match status as usize {
        sys::PB2_CAST_NONE => Ok(None),
        sys::PB2_CAST_CONVERGED => Ok(Some(hit(ShapeCastStatus::Converged))),
        sys::PB2_CAST_PENETRATING => Ok(Some(hit(ShapeCastStatus::PenetratingOrWithinTargetDist))),
        _ => Err(Unsupported),
    }
}

/// status 0 -> `Ok(None)`, 1 -> `Ok(Some(contact))`, anything else -> `Err(Unsupported)`: 2 is an unknown shape; 3 (reserved: an
/// EPA polytope beyond the device's overflow arena, which the reference's iteration cap does not reach) would make the chain
/// re-run that one pair on `DefaultQueryDispatcher`.
fn contact_of_status(status: u8, c: &sys::pb2_contact) -> Result<Option<Contact>, Unsupported> {
    match status {
        0 => Ok(None),
        1 => Ok(Some(Contact::new(
            Point::new(c.point1[0], c.point1[1], c.point1[2]),
            Point::new(c.point2[0], c.point2[1], c.point2[2]),
            nalgebra::Unit::new_unchecked(Vector::new(c.normal1[0], c.normal1[1], c.normal1[2])),
            nalgebra::Unit::new_unchecked(Vector::new(c.normal2[0], c.normal2[1], c.normal2[2])),
            c.dist,
        ))),
        _ => Err(Unsupported),
    }
}

/// `query::contact(pos1, g1, pos2, g2, prediction)` — contact_shape_shape.rs:123, for n pairs of table shapes.
pub fn contact_batch(table: &ShapeTable<'_>, shape1: &[u32], pos1: &[Isometry<Real>], shape2: &[u32], pos2: &[Isometry<Real>], prediction: Real)
                     -> Result<Vec<Result<Option<Contact>, Unsupported>>, Error> {
    let n = shape1.len();
    let (p1, p2): (Vec<[f32; 7]>, Vec<[f32; 7]>) = (pos1.iter().map(iso7).collect(), pos2.iter().map(iso7).collect());
    let (mut out, mut status) = (vec![sys::pb2_contact::default(); n], vec![0u8; n]);
    let h = table.h;
    table.c.call(|ctx| unsafe {
        sys::pb2_contact_batch(ctx, h, shape1.as_ptr(), shape2.as_ptr(), p1.as_ptr() as *const f32, p2.as_ptr() as *const f32, prediction, n as u32, out.as_mut_ptr(),
                               status.as_mut_ptr(), core::ptr::null_mut(), sys::PB2_MEM_HOST)
    })?;
    Ok((0..n).map(|k| contact_of_status(status[k], &out[k])).collect())
}

// ------------------------------------------------------------------------------------------------ QueryDispatcher
/// Key of one per-pair query as the trait sees it: the two shape addresses and the exact bits of the relative pose and scalar.
#[derive(Clone, Copy, PartialEq, Eq, Hash)]
struct PairKey {
    g1: usize,
    g2: usize,
    pos12: [u32; 7],
    scalar: u32,
}

fn pair_key(pos12: &Isometry<Real>, g1: &dyn Shape, g2: &dyn Shape, scalar: Real) -> PairKey {
    let p = iso7(pos12);
    PairKey { g1: shape_key(g1), g2: shape_key(g2), pos12: p.map(f32::to_bits), scalar: scalar.to_bits() }
}

/// One registered query of a step: `(g1, g2, pos12)` exactly as the narrow phase will pass them to the trait method.
pub struct PairQuery<'a> {
    pub pos12: Isometry<Real>,
    pub g1: &'a dyn Shape,
    pub g2: &'a dyn Shape,
}

/// `QueryDispatcher` over the device kernels. Use it chained: `B200Dispatcher::new(table).chain(DefaultQueryDispatcher)`.
pub struct B200Dispatcher<'c> {
    table: ShapeTable<'c>,
    contacts: Mutex<HashMap<PairKey, Result<Option<Contact>, Unsupported>>>,
    distances: Mutex<HashMap<PairKey, Result<Real, Unsupported>>>,
    intersections: Mutex<HashMap<PairKey, Result<bool, Unsupported>>>,
}

impl<'c> B200Dispatcher<'c> {
    pub fn new(table: ShapeTable<'c>) -> Self {
        B200Dispatcher { table, contacts: Mutex::new(HashMap::new()), distances: Mutex::new(HashMap::new()), intersections: Mutex::new(HashMap::new()) }
    }

    /// Splits the queries into the ones both of whose shapes are in the table (with their ids) and the rest.
    fn resolve<'a>(&self, queries: &'a [PairQuery<'a>]) -> (Vec<usize>, Vec<u32>, Vec<u32>, Vec<[f32; 7]>) {
        let (mut which, mut s1, mut s2, mut p) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (i, q) in queries.iter().enumerate() {
            if let (Some(a), Some(b)) = (self.table.id_of(q.g1), self.table.id_of(q.g2)) {
                which.push(i);
                s1.push(a);
                s2.push(b);
                p.push(iso7(&q.pos12));
            }
        }
        (which, s1, s2, p)
    }

    /// Runs `QueryDispatcher::contact` for every registered pair in ONE device batch (`pb2_contact_batch_local`: relative pose in,
    /// contact in the shapes' local frames out — query_dispatcher.rs:430-436) and keeps the answers for the per-pair calls that
    /// follow. Call once per step before the narrow phase iterates its pairs; `clear()` afterwards.
    pub fn prepare_contacts(&self, queries: &[PairQuery<'_>], prediction: Real) -> Result<(), Error> {
        let (which, s1, s2, p) = self.resolve(queries);
        let n = which.len();
        let (mut out, mut status) = (vec![sys::pb2_contact::default(); n], vec![0u8; n]);
        let h = self.table.h;
        self.table.c.call(|ctx| unsafe {
            sys::pb2_contact_batch_local(ctx, h, s1.as_ptr(), s2.as_ptr(), p.as_ptr() as *const f32, prediction, n as u32, out.as_mut_ptr(), status.as_mut_ptr(),
                                         sys::PB2_MEM_HOST)
        })?;
        let mut map = self.contacts.lock().unwrap();
        for (k, &i) in which.iter().enumerate() {
            let q = &queries[i];
            map.insert(pair_key(&q.pos12, q.g1, q.g2, prediction), contact_of_status(status[k], &out[k]));
        }
        Ok(())
    }

    /// Same for `distance` (default_query_dispatcher.rs:177-236) and `intersection_test` (:104-175). Both are invariant under a
    /// common rigid motion, so the world-pose entry points are called with (identity, pos12): `identity.inv_mul(pos12)` is
    /// `pos12` bit for bit.
    pub fn prepare_distances(&self, queries: &[PairQuery<'_>]) -> Result<(), Error> {
        let (which, s1, s2, p) = self.resolve(queries);
        let n = which.len();
        let ident = vec![IDENTITY7; n];
        let (mut dist, mut hit, mut st_d, mut st_h) = (vec![0f32; n], vec![0u8; n], vec![0u8; n], vec![0u8; n]);
        let h = self.table.h;
        self.table.c.call(|ctx| unsafe {
            sys::pb2_distance_batch(ctx, h, s1.as_ptr(), s2.as_ptr(), ident.as_ptr() as *const f32, p.as_ptr() as *const f32, n as u32, dist.as_mut_ptr(), st_d.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        self.table.c.call(|ctx| unsafe {
            sys::pb2_intersection_test_batch(ctx, h, s1.as_ptr(), s2.as_ptr(), ident.as_ptr() as *const f32, p.as_ptr() as *const f32, n as u32, hit.as_mut_ptr(), st_h.as_mut_ptr(), sys::PB2_MEM_HOST)
        })?;
        let (mut dm, mut im) = (self.distances.lock().unwrap(), self.intersections.lock().unwrap());
        for (k, &i) in which.iter().enumerate() {
            let q = &queries[i];
            let key = pair_key(&q.pos12, q.g1, q.g2, 0.0);
            dm.insert(key, if st_d[k] == 0 { Ok(dist[k]) } else { Err(Unsupported) });
            im.insert(key, if st_h[k] == 0 { Ok(hit[k] != 0) } else { Err(Unsupported) });
        }
        Ok(())
    }

    /// Forgets the prepared answers (end of the step).
    pub fn clear(&self) {
        self.contacts.lock().unwrap().clear();
        self.distances.lock().unwrap().clear();
        self.intersections.lock().unwrap().clear();
    }

    fn one<'a>(pos12: &Isometry<Real>, g1: &'a dyn Shape, g2: &'a dyn Shape) -> [PairQuery<'a>; 1] {
        [PairQuery { pos12: *pos12, g1, g2 }]
    }
}

impl QueryDispatcher for B200Dispatcher<'_> {
    fn intersection_test(&self, pos12: &Isometry<Real>, g1: &dyn Shape, g2: &dyn Shape) -> Result<bool, Unsupported> {
        let key = pair_key(pos12, g1, g2, 0.0);
        if let Some(r) = self.intersections.lock().unwrap().get(&key) {
            return *r;
        }
        if self.table.id_of(g1).is_none() || self.table.id_of(g2).is_none() {
            return Err(Unsupported);
        }
        self.prepare_distances(&Self::one(pos12, g1, g2)).map_err(|_| Unsupported)?;   // a batch of one: still the device path
        self.intersections.lock().unwrap().get(&key).copied().unwrap_or(Err(Unsupported))
    }

    fn distance(&self, pos12: &Isometry<Real>, g1: &dyn Shape, g2: &dyn Shape) -> Result<Real, Unsupported> {
        let key = pair_key(pos12, g1, g2, 0.0);
        if let Some(r) = self.distances.lock().unwrap().get(&key) {
            return *r;
        }
        if self.table.id_of(g1).is_none() || self.table.id_of(g2).is_none() {
            return Err(Unsupported);
        }
        self.prepare_distances(&Self::one(pos12, g1, g2)).map_err(|_| Unsupported)?;
        self.distances.lock().unwrap().get(&key).copied().unwrap_or(Err(Unsupported))
    }

    fn contact(&self, pos12: &Isometry<Real>, g1: &dyn Shape, g2: &dyn Shape, prediction: Real) -> Result<Option<Contact>, Unsupported> {
        let key = pair_key(pos12, g1, g2, prediction);
        if let Some(r) = self.contacts.lock().unwrap().get(&key) {
            return *r;
        }
        if self.table.id_of(g1).is_none() || self.table.id_of(g2).is_none() {
            return Err(Unsupported);   // e.g. a Capsule: the chain hands the pair to DefaultQueryDispatcher
        }
        self.prepare_contacts(&Self::one(pos12, g1, g2), prediction).map_err(|_| Unsupported)?;
        self.contacts.lock().unwrap().get(&key).copied().unwrap_or(Err(Unsupported))
    }

    /// `pb2_closest_points_batch` answers in world space; called with (identity, pos12) the second point comes back in shape 1's
    /// frame and is moved to shape 2's with `pos12.inverse_transform_point`, one rounding away from the reference's local value.
    fn closest_points(&self, pos12: &Isometry<Real>, g1: &dyn Shape, g2: &dyn Shape, max_dist: Real) -> Result<ClosestPoints, Unsupported> {
        let (Some(a), Some(b)) = (self.table.id_of(g1), self.table.id_of(g2)) else { return Err(Unsupported) };
        let p = iso7(pos12);
        let (mut pts, mut kind, mut status) = ([0f32; 6], 0u8, 0u8);
        let h = self.table.h;
        self.table.c.call(|ctx| unsafe {
            sys::pb2_closest_points_batch(ctx, h, &a, &b, IDENTITY7.as_ptr(), p.as_ptr(), max_dist, 1, pts.as_mut_ptr(), &mut kind, &mut status, sys::PB2_MEM_HOST)
        }).map_err(|_| Unsupported)?;
        if status != 1 {
            return Err(Unsupported);
        }
        Ok(match kind {
            0 => ClosestPoints::Disjoint,
            1 => ClosestPoints::WithinMargin(Point::new(pts[0], pts[1], pts[2]), pos12.inverse_transform_point(&Point::new(pts[3], pts[4], pts[5]))),
            _ => ClosestPoints::Intersecting,
        })
    }

    /// `cast_shapes(pos12, local_vel12, g1, g2, options)` — query_dispatcher.rs:445: the world entry point with (identity, 0) for
    /// shape 1 reproduces `pos12` and `local_vel12` exactly (shape_cast.rs:262-266); the hit is already in local frames.
    fn cast_shapes(&self, pos12: &Isometry<Real>, local_vel12: &Vector<Real>, g1: &dyn Shape, g2: &dyn Shape, options: ShapeCastOptions)
                   -> Result<Option<ShapeCastHit>, Unsupported> {
        let (Some(a), Some(b)) = (self.table.id_of(g1), self.table.id_of(g2)) else { return Err(Unsupported) };
        let (p, v1, v2) = (iso7(pos12), [0f32; 3], [local_vel12.x, local_vel12.y, local_vel12.z]);
        let (mut out, mut status) = ([0f32; 13], 0u8);
        let h = self.table.h;
        self.table.c.call(|ctx| unsafe {
            sys::pb2_cast_shapes_batch(ctx, h, &a, &b, IDENTITY7.as_ptr(), v1.as_ptr(), p.as_ptr(), v2.as_ptr(), options.max_time_of_impact, options.target_distance,
                                       options.stop_at_penetration as i32, options.compute_impact_geometry_on_penetration as i32, 1, out.as_mut_ptr(), &mut status,
                                       sys::PB2_MEM_HOST)
        }).map_err(|_| Unsupported)?;
        cast_of_status(status, &out)
    }

    /// Not on the device path (SURVEY §8 f3 lists it as open): always `Unsupported`, i.e. `DefaultQueryDispatcher` in a chain.
    fn cast_shapes_nonlinear(&self, _motion1: &NonlinearRigidMotion, _g1: &dyn Shape, _motion2: &NonlinearRigidMotion, _g2: &dyn Shape, _start_time: Real,
                             _end_time: Real, _stop_at_penetration: bool) -> Result<Option<ShapeCastHit>, Unsupported> {
        Err(Unsupported)
    }
}
