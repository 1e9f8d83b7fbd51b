"""parry_b200 — B200-native (sm_100a CUDA) implementation of parry3d's data-parallel query hot path, behind a C ABI
(include/parry_b200.h). This package is the thin host-side mirror of the reference's Rust API for that path
(`Bvh`, `TriMesh`/`RayCast`, `query::contact`), used by the parity tests and bench.py. No CPU fallback exists."""
from ._ffi import Pb2Error, Unsupported, INVALID_U32, lib  # noqa: F401
from .host import Context, Comm, Bvh, BvhBuildStrategy, TriMesh, Shapes, Compounds, Ball, Cuboid, ConvexPolyhedron, contact, contact_local, contact_compact, contact_pairs_compact, distance, intersection_test, cast_shapes, ShapeCastOptions, contact_manifolds, closest_points, manifolds_try_update, contact_manifolds_update  # noqa: F401
