"""ctypes binding of the C ABI declared in include/parry_b200.h (the same symbols a Rust `extern "C"` block binds,
see INTEGRATION.md). Fails loudly when the CUDA library is missing: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB2_LIB_PATH") or os.path.join(HERE, "libparry_b200.so")   # PB2_LIB_PATH: A/B builds of the same library

PB2_OK, PB2_ERR_INVALID, PB2_ERR_CUDA, PB2_ERR_OVERFLOW, PB2_ERR_UNSUPPORTED, PB2_ERR_DEPTH = 0, -1, -2, -3, -4, -5
MEM_HOST, MEM_DEVICE = 0, 1
INVALID_U32 = 0xFFFFFFFF

_lib = None

c_void_p, c_int, c_u32, c_u64, c_float = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
P = C.c_void_p  # every data pointer is passed as an address

# name -> (restype, argtypes); kept in sync with include/parry_b200.h (tests check every symbol is exported)
SIGNATURES = {
    "pb2_version": (c_int, []),
    "pb2_device_count": (c_int, []),
    "pb2_ctx_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "pb2_ctx_create_on_stream": (c_int, [c_int, c_void_p, C.POINTER(c_void_p)]),
    "pb2_ctx_destroy": (c_int, [c_void_p]),
    "pb2_ctx_synchronize": (c_int, [c_void_p]),
    "pb2_ctx_stream": (c_void_p, [c_void_p]),
    "pb2_last_error": (C.c_char_p, [c_void_p]),
    "pb2_ctx_launch_count": (c_u64, [c_void_p]),
    "pb2_trimesh_traversal_bytes": (c_u64, [c_void_p]),
    "pb2_ctx_enable_phase_timing": (c_int, [c_void_p, c_int]),
    "pb2_contact_phase_times": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pb2_bvh_build": (c_int, [c_void_p, P, c_u32, c_int, c_int, C.POINTER(c_void_p)]),
    "pb2_bvh_destroy": (c_int, [c_void_p, c_void_p]),
    "pb2_bvh_self_pairs_shard": (c_int, [c_void_p, c_void_p, c_int, c_u32, c_u32, P, c_u64, C.POINTER(c_u64), c_int]),
    "pb2_comm_unique_id": (c_int, [P]),
    "pb2_comm_create": (c_int, [c_void_p, P, c_int, c_int, C.POINTER(c_void_p)]),
    "pb2_comm_destroy": (c_int, [c_void_p]),
    "pb2_comm_rank": (c_int, [c_void_p]),
    "pb2_comm_size": (c_int, [c_void_p]),
    "pb2_comm_allgather": (c_int, [c_void_p, P, P, c_u64]),
    "pb2_comm_allgather_counts": (c_int, [c_void_p, c_u64, P]),
    "pb2_comm_allgatherv": (c_int, [c_void_p, P, c_u64, c_u32, P, c_u64, P, C.POINTER(c_u64)]),
    "pb2_comm_barrier": (c_int, [c_void_p]),
    "pb2_comm_peer_alloc": (c_int, [c_void_p, c_u64, P]),
    "pb2_bvh_leaf_count": (c_u32, [c_void_p]),
    "pb2_bvh_node_count": (c_u32, [c_void_p]),
    "pb2_bvh_update_leaves": (c_int, [c_void_p, c_void_p, P, P, c_u32, c_float, c_int]),
    "pb2_bvh_refit": (c_int, [c_void_p, c_void_p]),
    "pb2_bvh_remove_leaves": (c_int, [c_void_p, c_void_p, P, c_u32, c_int]),
    "pb2_bvh_resize": (c_int, [c_void_p, c_void_p, c_u32]),
    "pb2_bvh_rebuild": (c_int, [c_void_p, c_void_p, c_int]),
    "pb2_bvh_download": (c_int, [c_void_p, c_void_p, P, P, P, c_int]),
    "pb2_bvh_root_aabb": (c_int, [c_void_p, c_void_p, P]),
    "pb2_bvh_intersect_aabbs": (c_int, [c_void_p, c_void_p, P, c_u32, P, P, c_u64, C.POINTER(c_u64), c_int]),
    "pb2_bvh_self_pairs": (c_int, [c_void_p, c_void_p, c_int, P, c_u64, C.POINTER(c_u64), c_int]),
    "pb2_bvh_leaf_pairs": (c_int, [c_void_p, c_void_p, c_void_p, P, c_u64, C.POINTER(c_u64), c_int]),
    "pb2_trimesh_create": (c_int, [c_void_p, P, c_u32, P, c_u32, c_int, C.POINTER(c_void_p)]),
    "pb2_trimesh_destroy": (c_int, [c_void_p, c_void_p]),
    "pb2_trimesh_bvh": (c_void_p, [c_void_p]),
    "pb2_trimesh_cast_rays": (c_int, [c_void_p, c_void_p, P, P, c_u32, c_float, c_int, P, P, P, P, c_int]),
    "pb2_trimesh_cast_rays_with_culling": (c_int, [c_void_p, c_void_p, P, P, c_u32, c_float, c_int, P, P, P, P, c_int]),
    "pb2_shapes_create": (c_int, [c_void_p, P, P, c_u32, P, c_u32, C.POINTER(c_void_p)]),
    "pb2_shapes_destroy": (c_int, [c_void_p, c_void_p]),
    "pb2_shapes_compute_aabbs": (c_int, [c_void_p, c_void_p, P, P, c_u32, P, c_int]),
    "pb2_bvh_cast_rays_shapes": (c_int, [c_void_p, c_void_p, c_void_p, P, P, P, c_u32, c_float, c_int, P, P, P, P, c_int]),
    "pb2_bvh_project_points_shapes": (c_int, [c_void_p, c_void_p, c_void_p, P, P, P, c_u32, c_float, c_int, P, P, P, P, c_int]),
    "pb2_contact_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_float, c_u32, P, P, C.POINTER(c_u64), c_int]),
    "pb2_contact_batch_local": (c_int, [c_void_p, c_void_p, P, P, P, c_float, c_u32, P, P, c_int]),
    "pb2_contact_batch_compact": (c_int, [c_void_p, c_void_p, P, P, P, P, c_float, c_u32, P, P, c_u64, C.POINTER(c_u64), c_int]),
    "pb2_trimesh_cast_rays_allgather": (c_int, [c_void_p, c_void_p, P, P, c_u32, c_float, P, P, c_int, c_int, c_u64, c_int]),
    "pb2_trimesh_project_points": (c_int, [c_void_p, c_void_p, P, P, c_u32, c_int, P, P, P, c_int]),
    "pb2_trimesh_contact_shapes": (c_int, [c_void_p, c_void_p, P, c_void_p, P, P, c_u32, c_float, P, P, P, c_int]),
    "pb2_distance_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_u32, P, P, c_int]),
    "pb2_compounds_create": (c_int, [c_void_p, c_void_p, P, P, c_u32, P, P, c_u32, C.POINTER(c_void_p)]),
    "pb2_compounds_destroy": (c_int, [c_void_p, c_void_p]),
    "pb2_compound_contact_shapes": (c_int, [c_void_p, c_void_p, P, P, P, P, c_u32, c_float, c_int, P, P, P, c_int]),
    "pb2_shapes_set_hull_topology": (c_int, [c_void_p, c_void_p, P, P, P, P, P, c_u32, P, P, c_u32]),
    "pb2_shapes_set_hull_vertex_topology": (c_int, [c_void_p, c_void_p, P, P, P, P, c_u32, P, P, c_u32]),
    "pb2_contact_manifolds_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_float, c_u32, c_u32, P, P, P, P, c_int]),
    "pb2_compound_contact_compounds": (c_int, [c_void_p, c_void_p, P, P, P, P, c_u32, c_float, P, P, P, c_int]),
    "pb2_compound_contact_trimesh": (c_int, [c_void_p, c_void_p, P, P, c_void_p, P, c_u32, c_float, c_int, P, P, P, c_int]),
    "pb2_manifolds_try_update": (c_int, [c_void_p, P, P, c_u32, c_u32, c_float, c_float, P, P, P, P, c_int]),
    "pb2_contact_manifolds_update_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_float, c_u32, c_u32, P, P, P, P, P, P, c_int]),
    "pb2_closest_points_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_float, c_u32, P, P, P, c_int]),
    "pb2_cast_shapes_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, P, P, c_float, c_float, c_int, c_int, c_u32, P, P, c_int]),
    "pb2_trimesh_cast_shapes": (c_int, [c_void_p, c_void_p, P, P, c_void_p, P, P, P, c_int, c_float, c_float, c_int, c_int, c_u32, P, P, P, c_int]),
    "pb2_trimesh_cast_trimesh": (c_int, [c_void_p, c_void_p, P, P, c_void_p, P, P, c_float, c_float, c_int, c_int, c_u32, P, P, P, c_int]),
    "pb2_trimesh_distance_shapes": (c_int, [c_void_p, c_void_p, P, c_void_p, P, P, c_int, c_u32, P, P, P, c_int]),
    "pb2_intersection_test_batch": (c_int, [c_void_p, c_void_p, P, P, P, P, c_u32, P, P, c_int]),
    "pb2_contact_pairs_compact": (c_int, [c_void_p, c_void_p, P, P, c_u32, P, c_u32, c_float, P, P, c_u64, C.POINTER(c_u64), c_int]),
}


class Pb2Error(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__("parry_b200 status %d: %s" % (status, msg))
        self.status = status


class Unsupported(Pb2Error):
    """query::Unsupported (query/error.rs)."""


def lib():
    """Loads libparry_b200.so (building is __graft_entry__.build()'s job). Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "parry_b200: %s is missing — run `python -m parry_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l
