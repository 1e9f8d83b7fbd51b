/* parry_b200 — C ABI of the B200-native (sm_100a) implementation of parry3d's data-parallel query hot path.
 *
 * parry3d (reference, 100 % Rust) has no FFI of its own; the boundary it exposes for this path is its public
 * Rust API. Every entry point below is the *batched* form of one reference call and cites the reference
 * interface it replaces (paths relative to the parry checkout). INTEGRATION.md shows the Rust `extern "C"`
 * bindings + the shim types (`Bvh`, `TriMesh: RayCast`, `B200Dispatcher: QueryDispatcher`) built on them.
 *
 * Conventions
 *  - All functions return PB2_OK (0) or a negative pb2_status. Nothing throws / panics across the boundary.
 *  - `mem` says where EVERY data pointer of that call lives: PB2_MEM_HOST (the library stages through HBM and
 *    synchronises before returning) or PB2_MEM_DEVICE (pointers are device pointers on ctx's device; the call
 *    is enqueued on ctx's stream and returns without synchronising).
 *  - Layouts are the reference's `#[repr(C)]` ones: Aabb = {mins[3], maxs[3]} (24 B, bounding_volume/aabb.rs:110),
 *    Ray = {origin[3], dir[3]} (24 B, query/ray/ray.rs:74-88), Isometry3<f32> = {qi,qj,qk,qw, tx,ty,tz} (28 B,
 *    nalgebra field order), BvhNodeWide = 64 B (partitioning/bvh/bvh_tree.rs:263-266).
 *  - Output buffers are caller-owned with a capacity; the true count is always reported so overflow is
 *    detectable (PB2_ERR_OVERFLOW is returned, the first `cap` entries are valid).
 *  - A pb2_ctx is bound to one CUDA device + one stream and is not re-entrant; distinct ctxs are independent
 *    (mirrors `&self` readers / `&mut self` writers of the Rust API).
 *  - There is NO CPU fallback: without a CUDA device pb2_ctx_create fails with PB2_ERR_CUDA.
 */
#ifndef PARRY_B200_H
#define PARRY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pb2_status {
    PB2_OK = 0,
    PB2_ERR_INVALID = -1,     /* bad argument */
    PB2_ERR_CUDA = -2,        /* CUDA runtime error; see pb2_last_error */
    PB2_ERR_OVERFLOW = -3,    /* output capacity too small; *count holds the required size */
    PB2_ERR_UNSUPPORTED = -4, /* mirrors query::Unsupported (query/error.rs) */
    PB2_ERR_DEPTH = -5        /* a tree walk ran out of its fixed per-thread stack (tree deeper than 96 levels; impossible for the
                                 Morton-linked trees): results since the last synchronisation are incomplete.
                                 Sticky: reported by the next synchronising call or by pb2_ctx_synchronize, then cleared. */
} pb2_status;

typedef enum pb2_mem { PB2_MEM_HOST = 0, PB2_MEM_DEVICE = 1 } pb2_mem;

/* BvhBuildStrategy (partitioning/bvh/bvh_tree.rs:58-78). Both builders sort the leaves by 63-bit Morton code on the device.
 *  - PB2_BUILD_BINNED (the reference's default, bvh_binned_build.rs:39-176: sequential top-down 8-bin SAH): replaced by a plain
 *    Karras LBVH link of the sorted leaves (no SAH pass) — the per-frame builder, ~1 ms per million leaves.
 *  - PB2_BUILD_PLOC (bvh_ploc_build.rs:10-94): the same algorithm on the GPU — rounds of radius-16 nearest-neighbour search by
 *    half-area of the merged box and merging of mutual pairs; links the reference's own topology from the same sorted leaves
 *    (tests/test_hostcheck.py). Thousands of identical boxes stall the clustering; the builder then falls back to the LBVH link.
 * Query results never depend on the strategy. */
typedef enum pb2_build_strategy { PB2_BUILD_BINNED = 0, PB2_BUILD_PLOC = 1 } pb2_build_strategy;

/* Shape kinds known to the typed-leaf / contact kernels (shape/shape.rs ShapeType subset on the hot path). */
typedef enum pb2_shape_kind { PB2_SHAPE_BALL = 0, PB2_SHAPE_CUBOID = 1, PB2_SHAPE_CONVEX = 2 } pb2_shape_kind;

typedef struct pb2_ctx pb2_ctx;
typedef struct pb2_bvh pb2_bvh;
typedef struct pb2_trimesh pb2_trimesh;
typedef struct pb2_shapes pb2_shapes;
typedef struct pb2_compounds pb2_compounds;
typedef struct pb2_comm pb2_comm;

#define PB2_INVALID_U32 0xffffffffu

/* Contact (query/contact/contact.rs:71-105): point1, point2, normal1, normal2, dist = 13 f32 = 52 B. */
typedef struct pb2_contact {
    float point1[3];
    float point2[3];
    float normal1[3];
    float normal2[3];
    float dist;
} pb2_contact;

/* ------------------------------------------------------------------ context */
int pb2_version(void);
/* Number of visible CUDA devices (0 => nothing in this library can run). */
int pb2_device_count(void);
/* Creates a context on `device` with its own non-blocking stream. */
int pb2_ctx_create(int device, pb2_ctx** out);
/* Same, but enqueue on a caller-owned cudaStream_t (passed as void*). */
int pb2_ctx_create_on_stream(int device, void* cuda_stream, pb2_ctx** out);
int pb2_ctx_destroy(pb2_ctx* ctx);
int pb2_ctx_synchronize(pb2_ctx* ctx);
/* The cudaStream_t the context enqueues on (so callers can record events on it). */
void* pb2_ctx_stream(pb2_ctx* ctx);
const char* pb2_last_error(pb2_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t pb2_ctx_launch_count(pb2_ctx* ctx);
/* Measurement aid (no reference counterpart): with phase timing on, every contact-family call records CUDA events on the context's
 * stream around its GJK kernel, its EPA kernel and its finishing kernel; pb2_contact_phase_times waits for the last such call and
 * returns the three durations in milliseconds and the number of pairs GJK handed to EPA. */
int pb2_ctx_enable_phase_timing(pb2_ctx* ctx, int on);
int pb2_contact_phase_times(pb2_ctx* ctx, float* gjk_ms, float* epa_ms, float* finish_ms, uint64_t* epa_runs);

/* ------------------------------------------------------------------ Bvh (partitioning/bvh) */
/* Bvh::from_leaves(strategy, &[Aabb]) — bvh_tree.rs:1835 (leaf id = index). n may be 0, 1, 2 (special-cased like
 * bvh_tree.rs:1914-1932). */
int pb2_bvh_build(pb2_ctx* ctx, const float* aabbs /* n x 6 */, uint32_t n, int strategy, int mem, pb2_bvh** out);
int pb2_bvh_destroy(pb2_ctx* ctx, pb2_bvh* bvh);
/* Bvh::leaf_count — bvh_tree.rs:2295 */
uint32_t pb2_bvh_leaf_count(const pb2_bvh* bvh);
/* number of BvhNodeWide entries (`nodes.len()`) */
uint32_t pb2_bvh_node_count(const pb2_bvh* bvh);
/* Bvh::insert_or_update_partially(aabb, leaf_index, change_detection_margin) for n existing leaves —
 * bvh_insert.rs:209-231. ids == NULL means ids[i] = i. */
int pb2_bvh_update_leaves(pb2_ctx* ctx, pb2_bvh* bvh, const uint32_t* ids, const float* aabbs, uint32_t n, float margin, int mem);
/* Bvh::refit(&mut workspace) — bvh_refit.rs:170-180 (internal AABBs, leaf counts, change-flag resolution). */
/* Bvh::remove(leaf_index) (bvh_tree.rs:2360-2427), batched: the leaves stop taking part in every query (their slot keeps
 * Aabb::new_invalid()) and the ancestor boxes are re-fitted. Unknown ids are ignored, like the reference. */
int pb2_bvh_remove_leaves(pb2_ctx* ctx, pb2_bvh* bvh, const uint32_t* ids, uint32_t n, int mem);
/* Grows the leaf index space to new_n leaves; new indices start removed. Bvh::insert(aabb, leaf_index)
 * (bvh_insert.rs:126-197) for a new index = pb2_bvh_resize + pb2_bvh_update_leaves + pb2_bvh_rebuild: structural edits
 * are whole-tree rebuilds on the GPU. */
int pb2_bvh_resize(pb2_ctx* ctx, pb2_bvh* bvh, uint32_t new_n);
int pb2_bvh_refit(pb2_ctx* ctx, pb2_bvh* bvh);
/* Bvh::rebuild(&mut workspace, strategy) — bvh_binned_build.rs:11 (same leaves, new topology). */
int pb2_bvh_rebuild(pb2_ctx* ctx, pb2_bvh* bvh, int strategy);
/* Copies `nodes` out in the reference's BvhNodeWide layout (root at index 0; leaf `children` = leaf id;
 * `data` = leaf_count | change flags) so the Rust `Bvh` stays inspectable. parents / leaf_node_indices are
 * BvhNodeIndex values ((node << 1) | is_right, bvh_tree.rs:1284-1420) as u32; either may be NULL. */
int pb2_bvh_download(pb2_ctx* ctx, const pb2_bvh* bvh, void* nodes64, uint32_t* parents, uint32_t* leaf_node_indices, int mem);
/* Bvh::root_aabb — bvh_tree.rs:1991 (host pointer, 6 floats). */
int pb2_bvh_root_aabb(pb2_ctx* ctx, const pb2_bvh* bvh, float* aabb6);

/* Bvh::intersect_aabb(&aabb) for m query boxes — bvh_queries.rs:203-205. Two-pass CSR output:
 * offsets[m+1] (exclusive scan of per-query counts), leaf_ids[cap] grouped by query (order inside a group is
 * unspecified, like the reference's iterator order is tree-dependent). *count = total hits. */
int pb2_bvh_intersect_aabbs(pb2_ctx* ctx, const pb2_bvh* bvh, const float* queries /* m x 6 */, uint32_t m,
                            uint32_t* offsets, uint32_t* leaf_ids, uint64_t cap, uint64_t* count, int mem);
/* Bvh::traverse_bvtt_single_tree::<CHANGE_DETECTION>(ws, f) — bvh_traverse_bvtt.rs:19-31. Emits each unordered
 * overlapping leaf pair exactly once as (min id, max id); order unspecified. `count` is a HOST pointer. */
int pb2_bvh_self_pairs(pb2_ctx* ctx, const pb2_bvh* bvh, int change_detection, uint32_t* pairs /* cap x 2 */,
                       uint64_t cap, uint64_t* count, int mem);
/* Shard of the same pair set for a broad phase split over `n_shards` GPUs with the Bvh replicated: only the
 * leaves at sorted positions p = shard (mod n_shards) walk the tree (interleaved ownership balances the "other leaf comes later"
 * filter). The union of the n_shards outputs is exactly pb2_bvh_self_pairs' set, each pair in exactly one shard. */
int pb2_bvh_self_pairs_shard(pb2_ctx* ctx, const pb2_bvh* bvh, int change_detection, uint32_t shard, uint32_t n_shards,
                             uint32_t* pairs /* cap x 2 */, uint64_t cap, uint64_t* count, int mem);
/* Bvh::leaf_pairs(&other, |a, b| a.intersects(b)) — bvh_traverse_bvtt.rs:210-316 (leaf of a, leaf of b). */
int pb2_bvh_leaf_pairs(pb2_ctx* ctx, const pb2_bvh* a, const pb2_bvh* b, uint32_t* pairs, uint64_t cap,
                       uint64_t* count, int mem);

/* ------------------------------------------------------------------ TriMesh + RayCast (shape/trimesh.rs, query/ray) */
/* TriMesh::new(vertices, indices) — shape/trimesh.rs:607 / rebuild_bvh :1159-1171. nt == 0 is PB2_ERR_INVALID
 * (TriMeshBuilderError::EmptyIndices, trimesh.rs:724). */
int pb2_trimesh_create(pb2_ctx* ctx, const float* vertices /* nv x 3 */, uint32_t nv, const uint32_t* indices /* nt x 3 */,
                       uint32_t nt, int mem, pb2_trimesh** out);
int pb2_trimesh_destroy(pb2_ctx* ctx, pb2_trimesh* mesh);
/* Bytes of the arrays a ray cast walks (wide nodes + pre-gathered triangles): what one launch has to read at least once when
 * every part of the mesh is hit. Measurement aid for roofline figures in this library's own layout. */
uint64_t pb2_trimesh_traversal_bytes(const pb2_trimesh* mesh);
/* TriMesh::bvh() — borrowed handle, owned by the mesh. */
const pb2_bvh* pb2_trimesh_bvh(const pb2_trimesh* mesh);
/* RayCast::cast_ray / cast_ray_and_get_normal for TriMesh over m rays — query/ray/ray.rs:381-411,
 * ray_trimesh.rs:8-36, ray_composite_shape.rs:20-62, bvh_traverse.rs:335-417, ray_aabb.rs:12-49,
 * ray_triangle.rs:49-152.  pose7 may be NULL (identity / cast_local_ray*).  Outputs per ray: toi (0 on miss),
 * tri = hit triangle index or PB2_INVALID_U32 (None).  normal (m x 3) and feature (m; FeatureId::Face(i) or
 * Face(i + nt) for back faces; PB2_INVALID_U32 on miss) may both be NULL => toi-only variant.
 * Ties: among bit-equal minimal toi the smallest triangle index wins (documented rule). */
int pb2_trimesh_cast_rays(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays /* m x 6 */,
                          uint32_t m, float max_toi, int solid, float* toi, uint32_t* tri, float* normal,
                          uint32_t* feature, int mem);
/* PointQuery::project_point(m, pt, solid) on a TriMesh, batched (query/point/point_query.rs:147-151 ->
 * point_composite_shape.rs:164-186,49-72: Bvh::find_best on the distance to the node boxes, point_triangle.rs at the
 * leaves). proj: m x 3 projected points (world space), inside[k]: PointProjection::is_inside, tri[k]: the triangle the
 * point was projected on (FeatureId::Face of project_point_and_get_feature); equal distances resolve to the smallest
 * triangle index. */
int pb2_trimesh_project_points(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* points /* m x 3 */, uint32_t m,
                               int solid, float* proj, uint8_t* inside, uint32_t* tri, int mem);

/* Multi-GPU form of pb2_trimesh_cast_rays for a range-split batch with the mesh replicated per GPU (the
 * path's only exchange is the all-gather of the fixed-size (toi, tri) hit records). Everything is device memory. The local
 * shard of m rays is traversed in `chunks` pieces; as soon as a piece is done its results — written straight into this
 * rank's gather buffers at element offset `elem_offset` — are pushed into the same place of every peer's buffers with
 * device-to-device copies over NVLink (copy engines), while the next piece is traversed. peer_toi[p] / peer_tri[p]: rank
 * p's gather buffers as mapped in THIS process (CUDA IPC / symmetric memory; entry `self` is the local one). Asynchronous
 * on the context's stream; the caller runs a cross-rank barrier before reading peers' slices. */
int pb2_trimesh_cast_rays_allgather(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays /* m x 6, device */,
                                    uint32_t m, float max_toi, void* const* peer_toi, void* const* peer_tri, int n_peers, int self,
                                    uint64_t elem_offset, int chunks);

/* ---- Multi-GPU exchange, one process per GPU, one pb2_comm per context; NCCL underneath (resolved with
 * dlopen at the first call: PB2_ERR_UNSUPPORTED when libnccl.so.2 is absent). The reference has no counterpart: its callers
 * parallelise with rayon on one host. Rendezvous: rank 0 calls pb2_comm_unique_id and the host carries the PB2_COMM_ID_BYTES bytes
 * to the other ranks by whatever transport it has; then every rank calls pb2_comm_create (collective). All data pointers are
 * device memory; everything is enqueued on the context's stream unless stated otherwise. */
#define PB2_COMM_ID_BYTES 128
int pb2_comm_unique_id(void* id128);
int pb2_comm_create(pb2_ctx* ctx, const void* id128, int rank, int nranks, pb2_comm** out);
int pb2_comm_destroy(pb2_comm* comm);
int pb2_comm_rank(const pb2_comm* comm);
int pb2_comm_size(const pb2_comm* comm);
/* Fixed-size all-gather (per-ray hit records): recv = nranks x bytes_per_rank, rank-major; in place when send == recv + rank * bytes. */
int pb2_comm_allgather(pb2_comm* comm, const void* send, void* recv, uint64_t bytes_per_rank);
/* All-gather of one count per rank ("compacted hit / pair counts"); `all` is a HOST array of nranks; synchronises the stream. */
int pb2_comm_allgather_counts(pb2_comm* comm, uint64_t mine, uint64_t* all);
/* Variable-size gather of compacted records (pairs, contacts): rank r contributes `count` elements of elem_bytes; recv receives
 * every rank's elements back to back in rank order, counts[r] (HOST) = elements of rank r, *total = their sum. PB2_ERR_OVERFLOW
 * (nothing written, *total valid) when total > cap_elems. */
int pb2_comm_allgatherv(pb2_comm* comm, const void* send, uint64_t count, uint32_t elem_bytes, void* recv, uint64_t cap_elems,
                        uint64_t* counts, uint64_t* total);
/* Cross-rank barrier on the stream. */
int pb2_comm_barrier(pb2_comm* comm);
/* Collective: every rank allocates `bytes` and maps all other ranks' allocations (CUDA IPC, one node): peers[r] is rank r's
 * buffer as addressable from THIS process (peers[rank] = the local one) — the peer pointers pb2_trimesh_cast_rays_allgather
 * pushes finished result ranges to. Freed by pb2_comm_destroy. */
int pb2_comm_peer_alloc(pb2_comm* comm, uint64_t bytes, void** peers /* nranks */);

/* TriMesh::cast_ray_with_culling / cast_local_ray_with_culling (ray_trimesh.rs:139-178, RayCullingMode :50-65): same
 * query, but a triangle is only considered when its scaled normal faces the ray the allowed way. Always the
 * `_and_get_normal` flavour in the reference; normal / feature may still be NULL here. */
#define PB2_CULL_IGNORE_BACKFACES 1
#define PB2_CULL_IGNORE_FRONTFACES 2
int pb2_trimesh_cast_rays_with_culling(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays /* m x 6 */,
                                       uint32_t m, float max_toi, int culling, float* toi, uint32_t* tri, float* normal,
                                       uint32_t* feature, int mem);

/* ------------------------------------------------------------------ typed shape tables */
/* A table of shapes referenced by index from ray / contact batches. kinds[n] (pb2_shape_kind), params[n x 4]:
 * ball {r,-,-,-}; cuboid {hx,hy,hz,-}; convex {first_point, num_points (as u32 bit patterns), -, -} into
 * `points` (np x 3, ConvexPolyhedron::points(), shape/convex_polyhedron.rs:172-185). Always HOST pointers. */
int pb2_shapes_create(pb2_ctx* ctx, const uint8_t* kinds, const float* params, uint32_t n, const float* points,
                      uint32_t np, pb2_shapes** out);
int pb2_shapes_destroy(pb2_ctx* ctx, pb2_shapes* shapes);

/* Face topology of the table's ConvexPolyhedron entries, as parry builds it when such a shape is created
 * (ConvexPolyhedron::from_convex_mesh, shape/convex_polyhedron.rs:390-637; accessors faces(), vertices_adj_to_face(),
 * edges_adj_to_face()): hull_face_first / hull_face_count per table entry (ignored for balls and cuboids) select the hull's
 * faces in face_normal (nf x 3), face_first / face_count (nf; ranges of the two adjacency arrays, rebased to the concatenated
 * arrays); vertices_adj_to_face holds vertex ids local to the hull, edges_adj_to_face the hull's edge ids. Only the pfm_pfm
 * contact-manifold arm needs it (PolygonalFeatureMap::local_support_feature, :959-991). HOST pointers; set at most once. */
int pb2_shapes_set_hull_topology(pb2_ctx* ctx, pb2_shapes* shapes, const uint32_t* hull_face_first, const uint32_t* hull_face_count,
                                 const float* face_normal, const uint32_t* face_first, const uint32_t* face_count, uint32_t nf,
                                 const uint32_t* vertices_adj_to_face, const uint32_t* edges_adj_to_face, uint32_t nadj);

/* Vertex side of the same topology, for ConvexPolyhedron::support_feature_id_toward (convex_polyhedron.rs:885-922; the feature id
 * of the ball-vs-hull manifold arm): vert_first / vert_count per point of the table (parry's Vertex::first_adj_face_or_edge /
 * num_adj_faces_or_edge, rebased to the concatenated arrays), faces_adj_to_vertex / edges_adj_to_vertex (ids local to the hull),
 * hull_edge_first per table entry into edge_dir (ne x 3, Edge::dir). Call after pb2_shapes_set_hull_topology. */
int pb2_shapes_set_hull_vertex_topology(pb2_ctx* ctx, pb2_shapes* shapes, const uint32_t* vert_first, const uint32_t* vert_count,
                                        const uint32_t* faces_adj_to_vertex, const uint32_t* edges_adj_to_vertex, uint32_t nadj,
                                        const uint32_t* hull_edge_first, const float* edge_dir, uint32_t ne);

/* Shape::compute_aabb(pos) for n colliders (shape/shape.rs:369; aabb_ball.rs:25, aabb_cuboid.rs:9-16,
 * aabb_convex_polyhedron.rs:8) -> aabbs (n x 6). */
int pb2_shapes_compute_aabbs(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape_ids, const float* poses7,
                             uint32_t n, float* aabbs, int mem);

/* Bvh::cast_ray with typed leaves (bvh_queries.rs:260-271 + RayCast for Ball ray_ball.rs:8-98 / Cuboid
 * ray_cuboid.rs:6-25 + ray_aabb.rs:52-92 + clip_aabb_line.rs:79-187 / ConvexPolyhedron ray_support_map.rs:19-72,163-181
 * + gjk.rs:519-534,660-795): leaf i of `bvh` is shape shape_ids[i] at poses7[i]. feature: Face(i) as i for ball / cuboid
 * leaves, PB2_FEATURE_UNKNOWN (FeatureId::Unknown) for convex leaves, 0xFFFFFFFF on a miss. normal/feature may be NULL. */
#define PB2_FEATURE_UNKNOWN 0xFFFFFFFEu
int pb2_bvh_cast_rays_shapes(pb2_ctx* ctx, const pb2_bvh* bvh, const pb2_shapes* shapes, const uint32_t* shape_ids,
                             const float* poses7, const float* rays, uint32_t m, float max_toi, int solid, float* toi,
                             uint32_t* leaf, float* normal, uint32_t* feature, int mem);

/* Bvh::project_point with typed leaves (bvh_queries.rs:213-227: find_best on Aabb::distance_to_local_point, point_aabb.rs:135-146;
 * leaf check PointQuery::project_point(pose, pt, solid), point_query.rs:147-151, of a Ball point_ball.rs:9-21 / Cuboid point_aabb.rs:9-60
 * / ConvexPolyhedron point_support_map.rs:17-52): leaf i of `bvh` is shape shape_ids[i] (NULL: shape i) at poses7[i]. proj: m x 3
 * world-space projections; inside[k] = PointProjection::is_inside; leaf[k] = the leaf projected on (equal distances: smallest index)
 * or 0xFFFFFFFF; status[k]: 0 nothing within max_distance, 1 found, 3 = solid == 0 and the point lies inside a ConvexPolyhedron
 * leaf (the reference runs EPA there): host. */
int pb2_bvh_project_points_shapes(pb2_ctx* ctx, const pb2_bvh* bvh, const pb2_shapes* shapes, const uint32_t* shape_ids, const float* poses7,
                                  const float* points /* m x 3 */, uint32_t m, float max_distance, int solid, float* proj, uint8_t* inside,
                                  uint32_t* leaf, uint8_t* status, int mem);

/* ------------------------------------------------------------------ query::contact (query/contact, gjk, epa) */
/* query::contact(pos1, g1, pos2, g2, prediction) for n pairs — contact_shape_shape.rs:123-138 through
 * DefaultQueryDispatcher::contact (default_query_dispatcher.rs:302-356). Pair k uses shapes shape1[k] /
 * shape2[k] at poses pos1[k] / pos2[k]. status[k]: 0 = Ok(None), 1 = Ok(Some(contact)) (out[k] valid),
 * 2 = Err(Unsupported).  *num_contacts (HOST pointer, may be NULL) = number of status==1. */
int pb2_contact_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                      const float* pos1 /* n x 7 */, const float* pos2 /* n x 7 */, float prediction, uint32_t n,
                      pb2_contact* out, uint8_t* status, uint64_t* num_contacts, int mem);
/* QueryDispatcher::contact(pos12, g1, g2, prediction) — query_dispatcher.rs:430-436, the trait method itself: pos12[k] is the pose
 * of shape 2 relative to shape 1 and the contact stays in the shapes' local frames (point1 / normal1 in shape 1's frame, point2 /
 * normal2 in shape 2's), which is what a dispatcher inside a QueryDispatcherChain has to return (query::contact applies
 * Contact::transform_by_mut afterwards, contact_shape_shape.rs:131-135). Same status codes as pb2_contact_batch. */
int pb2_contact_batch_local(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                            const float* pos12 /* n x 7 */, float prediction, uint32_t n, pb2_contact* out, uint8_t* status, int mem);
/* Compacted variant: writes only the Some(contact) records, each tagged with its pair index, through
 * warp-aggregated atomics (order unspecified). */
int pb2_contact_batch_compact(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                              const float* pos1, const float* pos2, float prediction, uint32_t n, pb2_contact* out,
                              uint32_t* pair_index, uint64_t cap, uint64_t* count, int mem);

/* query::distance(pos1, g1, pos2, g2) (query/distance/distance.rs:89-97) and query::intersection_test (query/intersection_test/
 * intersection_test.rs:88-96) for n pairs, through the same arms as DefaultQueryDispatcher (default_query_dispatcher.rs:104-236):
 * ball-ball closed forms, ball vs cuboid / hull by point projection, support-map pairs by GJK. status[k]: 0 = Ok (dist[k] /
 * hit[k] valid), 2 = Err(Unsupported) (unknown shape), 3 = cuboid-cuboid pair: the reference's SAT arm is not built here, the
 * host answers it (dispatcher chain). */
int pb2_distance_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1 /* n x 7 */,
                       const float* pos2 /* n x 7 */, uint32_t n, float* dist, uint8_t* status, int mem);
int pb2_intersection_test_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                                const float* pos2, uint32_t n, uint8_t* hit, uint8_t* status, int mem);

/* Narrow phase straight from a broad-phase pair list (BASELINE config 5: Bvh pair query feeding per-pair contacts; the
 * loop a caller such as rapier runs over the pairs reported by Bvh::traverse_bvtt_single_tree /
 * Bvh::leaf_pairs, bvh_traverse_bvtt.rs:19,210, calling query::contact, contact_shape_shape.rs:123, on each):
 * pair k = colliders (pairs[2k], pairs[2k+1]); collider i has shape collider_shape[i] at pose collider_pose[i].
 * Shapes and poses are read through the pair list inside the kernel (no gathered per-pair copies). Output as in
 * pb2_contact_batch_compact; pair_index[j] = k. Pairs naming a collider >= n_colliders are skipped. */
int pb2_contact_pairs_compact(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* collider_shape /* n_colliders */,
                              const float* collider_pose /* n_colliders x 7 */, uint32_t n_colliders,
                              const uint32_t* pairs /* n x 2 */, uint32_t n, float prediction, pb2_contact* out,
                              uint32_t* pair_index, uint64_t cap, uint64_t* count, int mem);

/* query::contact(mesh_pose, &TriMesh, poses7[k], shape k, prediction) for n shapes against ONE mesh: the composite arm of
 * DefaultQueryDispatcher::contact (default_query_dispatcher.rs:343-346) = contact_composite_shape_shape /
 * CompositeShapeRef::contact_with_shape (contact_composite_shape_shape.rs:14-61): the mesh Bvh is queried with
 * shape.compute_aabb(pose12).loosened(prediction), every reported triangle is dispatched as a shape::Triangle (ball:
 * contact_convex_polyhedron_ball on the triangle's point projection; cuboid / convex: GJK + EPA with the triangle's own
 * support map, shape/triangle.rs:697-716) and the contact with the smallest dist is kept. part[k] = winning triangle
 * (UINT32_MAX when status[k] != 1); equal dists resolve to the smallest triangle index. Contacts are in world space. */
int pb2_trimesh_contact_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const pb2_shapes* shapes,
                               const uint32_t* shape_ids /* n */, const float* poses7 /* n x 7 */, uint32_t n, float prediction,
                               pb2_contact* out, uint8_t* status, uint32_t* part, int mem);

/* Compound shapes (shape/compound.rs:113-144): compound c = parts comp_first[c] .. comp_first[c] + comp_count[c] - 1 of a part
 * table, part i = shape part_shape[i] of `shapes` at part_pose7[i] in the compound's frame. HOST pointers; `shapes` must
 * outlive the handle. Empty compounds are PB2_ERR_INVALID (Compound::new panics on them). */
int pb2_compounds_create(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* comp_first, const uint32_t* comp_count, uint32_t nc,
                         const uint32_t* part_shape, const float* part_pose7, uint32_t np, pb2_compounds** out);
int pb2_compounds_destroy(pb2_ctx* ctx, pb2_compounds* compounds);

/* query::contact with a Compound on one side, n pairs (composite arms of DefaultQueryDispatcher::contact,
 * default_query_dispatcher.rs:338-351 -> contact_composite_shape_shape.rs:14-76): pair k = compound compound_ids[k] at
 * compound_poses7[k] and shape shape_ids[k] at shape_poses7[k]. compound_second = 0: contact(compound, shape);
 * != 0: contact(shape, compound) (pose12.inverse() + Contact::flipped(), :63-76). out / status as pb2_contact_batch; part[k]
 * = index of the winning part inside its compound (equal dists: smallest index) or 0xFFFFFFFF. */
int pb2_compound_contact_shapes(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* compound_ids, const float* compound_poses7,
                                const uint32_t* shape_ids, const float* shape_poses7, uint32_t n, float prediction, int compound_second,
                                pb2_contact* out, uint8_t* status, uint32_t* part, int mem);

/* query::contact between two Compounds of the table, n pairs (default_query_dispatcher.rs:338-351, the composite arm nested through
 * contact_shape_composite_shape, contact_composite_shape_shape.rs:14-76; Compound::local_aabb compound.rs:120-127, Aabb::transform_by
 * aabb.rs:492-498): pair k = compound ids1[k] at poses1[k] vs compound ids2[k] at poses2[k]. out / status as pb2_contact_batch
 * (status 2: unknown compound id); parts: n x 2 = winning part of each compound (equal dists: smallest (part1, part2)) or
 * 0xFFFFFFFF. */
int pb2_compound_contact_compounds(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* ids1, const float* poses1 /* n x 7 */,
                                   const uint32_t* ids2, const float* poses2, uint32_t n, float prediction, pb2_contact* out, uint8_t* status,
                                   uint32_t* parts /* n x 2 */, int mem);

/* query::contact between a Compound and a TriMesh, n compounds against one mesh (default_query_dispatcher.rs:338-351: the composite arm
 * of shape 1, the other composite handled by the nested dispatch). mesh_first = 0: contact(compound_poses7[k], Compound compound_ids[k],
 * mesh_pose7, &TriMesh) — contact_composite_shape_shape.rs:12-48 over the parts (Bvh::root_aabb bvh_tree.rs:1991-1999 for the mesh's
 * box), :63-76 + :12-48 over the triangles; mesh_first != 0: contact(mesh_pose7, &TriMesh, compound_poses7[k], Compound) — over the
 * triangles first (Compound::local_aabb compound.rs:120-127), each against the compound's parts with the part as shape 1 of the leaf
 * problem. out / status as pb2_contact_batch (status 2: unknown compound id), contact sides in the argument order; parts: n x 2 =
 * {winning part, winning triangle} or 0xFFFFFFFF. Equal dists: the first part in part order, the smallest triangle index (the
 * reference keeps the first in its own tree's order). */
int pb2_compound_contact_trimesh(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* compound_ids, const float* compound_poses7 /* n x 7 */,
                                 const pb2_trimesh* mesh, const float* mesh_pose7, uint32_t n, float prediction, int mesh_first, pb2_contact* out,
                                 uint8_t* status, uint32_t* parts /* n x 2 */, int mem);

/* query::closest_points for n pairs (closest_points/closest_points_shape_shape.rs:220-231 -> default_query_dispatcher.rs:358-424:
 * closest_points_ball_ball.rs:7-36, closest_points_ball_convex_polyhedron.rs:7-44, closest_points_support_map_support_map.rs:8-69).
 * kind: 0 ClosestPoints::Disjoint, 1 WithinMargin (points[k] = p1, p2 in world space), 2 Intersecting. status: 1 ok, 2 unknown
 * shape id (3, host fallback, is reserved: no input of the three shape types produces it). */
int pb2_closest_points_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                             const float* pos1 /* n x 7 */, const float* pos2, float max_dist, uint32_t n, float* points /* n x 6 */,
                             uint8_t* kind, uint8_t* status, int mem);

/* QueryDispatcher::contact_manifolds for n pairs, first frame (empty incoming manifolds), pos12 =
 * pos1.inv_mul(pos2) (default_query_dispatcher.rs:629-835 -> contact_manifolds_ball_ball.rs:17-57,
 * contact_manifolds_convex_ball.rs:42-145, contact_manifolds_cuboid_cuboid.rs:19-107 + sat_cuboid_cuboid.rs +
 * polygonal_feature3d.rs:215-439). normals: n x 6 {ContactManifold::local_n1, local_n2}; counts: points per manifold;
 * points: n x max_points x 9 words {TrackedContact::local_p1 (3 f32), local_p2 (3 f32), dist (f32), fid1, fid2
 * (PackedFeatureId bits, u32)} in the reference's order. Pairs with a ConvexPolyhedron (against a Cuboid or another one) take
 * the pfm_pfm arm (contact_manifolds_pfm_pfm.rs:42-162: GJK/EPA contact, support faces along its normals, the same face
 * clipping, plus the witness pair as a feature-less point) once pb2_shapes_set_hull_topology has been called; Ball vs
 * ConvexPolyhedron (contact_manifolds_convex_ball.rs) additionally needs pb2_shapes_set_hull_vertex_topology. status: 0 ok,
 * 2 unsupported pair (a ConvexPolyhedron without topology — ball vs ConvexPolyhedron needs the vertex side too —, or an unknown shape id:
 * Err(Unsupported) / host), 3 host fallback (EPA arena overflow), 4 more than max_points contacts (two quads yield at most
 * 16; 9 is the observed max). */
int pb2_contact_manifolds_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                                const float* pos1 /* n x 7 */, const float* pos2, float prediction, uint32_t n, uint32_t max_points,
                                float* normals, uint32_t* counts, float* points, uint8_t* status, int mem);

/* ContactManifold::try_update_contacts_eps (query/contact_manifolds/contact_manifold.rs:662-699) on n manifolds in the layout of
 * pb2_contact_manifolds_batch, under the new pos12 = pos1.inv_mul(pos2): kept[k] = 1 when the manifold's normal turned by less
 * than acos(angle_dot_threshold) and no point drifted by more than sqrt(dist_sq_threshold) or switched between penetrating and
 * separated; dist and local_p1 of the visited points are refreshed in place (points: in / out). try_update_contacts (:652-658)
 * = thresholds COS_1_DEGREES (0.99984769515) and 1.0e-6. */
int pb2_manifolds_try_update(pb2_ctx* ctx, const float* pos1 /* n x 7 */, const float* pos2, uint32_t n, uint32_t max_points,
                             float angle_dot_threshold, float dist_sq_threshold, const float* normals, const uint32_t* counts,
                             float* points, uint8_t* kept, int mem);

/* QueryDispatcher::contact_manifolds called again with last frame's manifolds (PersistentQueryDispatcher::contact_manifold_convex_convex,
 * default_query_dispatcher.rs:760-831): normals / counts / points are in / out, in the layout of pb2_contact_manifolds_batch.
 * Cuboid-cuboid (contact_manifolds_cuboid_cuboid.rs:28) and pfm_pfm pairs (contact_manifolds_pfm_pfm.rs:63) whose manifold
 * passes try_update_contacts keep it (kept[k] = 1, status 0); every other pair is recomputed exactly as pb2_contact_manifolds_batch
 * would (the ball arms never keep). match (optional, n x max_points): for each new point the index of last frame's point
 * ContactManifold::match_contacts (contact_manifold.rs:761-770) would copy the ContactData from — the last one with both feature
 * ids equal — or -1; kept manifolds map onto themselves. One documented difference: the reference seeds the GJK of a pfm_pfm
 * recomputation with last frame's normal (contact_manifolds_pfm_pfm.rs:66); this path restarts from the default direction, so
 * those manifolds agree with the reference's within GJK's convergence tolerance rather than bit for bit. */
int pb2_contact_manifolds_update_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                                       const float* pos1 /* n x 7 */, const float* pos2, float prediction, uint32_t n, uint32_t max_points,
                                       float* normals, uint32_t* counts, float* points, uint8_t* status, uint8_t* kept, int32_t* match,
                                       int mem);

/* query::cast_shapes for n pairs (query/shape_cast/shape_cast.rs:268-286 -> DefaultQueryDispatcher::cast_shapes,
 * default_query_dispatcher.rs:434-515: ball-ball shape_cast_ball_ball.rs:10-69, every other Ball / Cuboid / ConvexPolyhedron
 * pair shape_cast_support_map_support_map.rs:11-69 + gjk::directional_distance gjk.rs:632-795). vel1 / vel2: n x 3
 * (world space). The four scalars are ShapeCastOptions (shape_cast.rs:196-243). out: n x 13 floats {witness1, witness2,
 * normal1, normal2, time_of_impact}, witness / normal i in the local frame of shape i as in ShapeCastHit; status:
 * PB2_CAST_*. */
#define PB2_CAST_NONE 0
#define PB2_CAST_CONVERGED 1       /* ShapeCastStatus::Converged */
#define PB2_CAST_PENETRATING 2     /* ShapeCastStatus::PenetratingOrWithinTargetDist */
#define PB2_CAST_UNSUPPORTED 3     /* shape id out of range */
#define PB2_CAST_NEEDS_HOST 4      /* the penetration contact overflowed the EPA arena (256 faces; seen with Ball support maps, whose
                                    * EPA runs converge slowly): run this pair on the host dispatcher */
int pb2_cast_shapes_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                          const float* pos1 /* n x 7 */, const float* vel1 /* n x 3 */, const float* pos2, const float* vel2,
                          float max_time_of_impact, float target_distance, int stop_at_penetration,
                          int compute_impact_geometry_on_penetration, uint32_t n, float* out /* n x 13 */, uint8_t* status, int mem);

/* query::cast_shapes with a TriMesh on one side, n queries against one mesh (the composite arms of DefaultQueryDispatcher::cast_shapes,
 * default_query_dispatcher.rs:498-515 -> shape_cast_composite_shape_shape.rs:14-105: Bvh::find_best over the mesh tree with
 * Minkowski-summed node boxes, every reached triangle cast like a support-map pair). mesh_second = 0:
 * cast_shapes(mesh_pose, mesh_vel, mesh, poses[k], vels[k], shape_ids[k]); != 0: the shape is shape 1 and the hit is
 * ShapeCastHit::swapped(). out / status as pb2_cast_shapes_batch (witness / normal of the mesh in the mesh's frame); part[k] = the
 * triangle that was hit or 0xFFFFFFFF. Equal times of impact resolve to the smallest triangle index (the reference keeps the
 * first in its own tree's order). stop_at_penetration = 0 is PB2_ERR_UNSUPPORTED (a leaf's answer would depend on its EPA contact). */
int pb2_trimesh_cast_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const float* mesh_vel3, const pb2_shapes* shapes,
                            const uint32_t* shape_ids, const float* poses7 /* n x 7 */, const float* vels3 /* n x 3 */, int mesh_second,
                            float max_time_of_impact, float target_distance, int stop_at_penetration,
                            int compute_impact_geometry_on_penetration, uint32_t n, float* out /* n x 13 */, uint8_t* status, uint32_t* part,
                            int mem);

/* query::cast_shapes between two TriMeshes, n pose / velocity pairs for the same two meshes (the nesting the reference's
 * crates/parry3d/tests/geometry/trimesh_trimesh_toi.rs exercises, issue #194: cast_shapes_composite_shape_shape over mesh 1, every
 * reached triangle through cast_shapes_shape_composite_shape over mesh 2, shape_cast_composite_shape_shape.rs:65-105). out as
 * pb2_cast_shapes_batch; parts: n x 2 = {triangle of mesh 1, triangle of mesh 2} or 0xFFFFFFFF. A winning pair that starts in touch
 * (time < 1e-5) takes its witnesses / normals from the triangle-triangle contact, as the reference does when
 * compute_impact_geometry_on_penetration is set. stop_at_penetration = 0 is PB2_ERR_UNSUPPORTED. */
int pb2_trimesh_cast_trimesh(pb2_ctx* ctx, const pb2_trimesh* mesh1, const float* pos1 /* n x 7 */, const float* vel1 /* n x 3 */,
                             const pb2_trimesh* mesh2, const float* pos2, const float* vel2, float max_time_of_impact, float target_distance,
                             int stop_at_penetration, int compute_impact_geometry_on_penetration, uint32_t n, float* out /* n x 13 */,
                             uint8_t* status, uint32_t* parts /* n x 2 */, int mem);

/* query::distance with a TriMesh on one side, n queries against one mesh (the composite arms of DefaultQueryDispatcher::distance,
 * default_query_dispatcher.rs:288-297 -> distance_composite_shape_shape.rs:13-77: Bvh::find_best with Aabb::distance_to_origin of the
 * Minkowski-summed node boxes, leaf = distance(triangle, shape)). mesh_second = 0: distance(mesh_pose, mesh, poses[k], shape k);
 * != 0: distance(poses[k], shape k, mesh_pose, mesh). dist[k] as the reference (0 when touching or penetrating; f32::MAX for a mesh
 * without live triangles); status[k]: 0 Ok, 2 unknown shape id; part[k] = the closest triangle (equal distances: smallest index). */
int pb2_trimesh_distance_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const pb2_shapes* shapes,
                                const uint32_t* shape_ids, const float* poses7 /* n x 7 */, int mesh_second, uint32_t n, float* dist,
                                uint8_t* status, uint32_t* part, int mem);

#ifdef __cplusplus
}
#endif
#endif /* PARRY_B200_H */
