"""TEST INFRASTRUCTURE. Face / edge topology of a convex polyhedron as parry builds it in
ConvexPolyhedron::from_convex_mesh (shape/convex_polyhedron.rs:390-637), restated in numpy on float32. In a real
deployment parry itself supplies these arrays (ConvexPolyhedron::faces(), edges(), vertices_adj_to_face(),
edges_adj_to_face()); here the triangle list comes from scipy's Qhull instead of parry's own convex_hull3, so face /
edge numbering is self-consistent between the oracle and the GPU but not parry's."""
import numpy as np

EPS = np.float32(1.1920929e-7)
F = np.float32


def _unit(v):
    n = F(np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))))
    return (v / n).astype(np.float32), n


def hull_triangles(points):
    """Outward-oriented (counter-clockwise seen from outside) triangles of the convex hull of `points` (n, 3)."""
    from scipy.spatial import ConvexHull
    h = ConvexHull(np.asarray(points, dtype=np.float64))
    tris = h.simplices.copy()
    p = np.asarray(points, dtype=np.float64)
    for k, (a, b, c) in enumerate(tris):
        n = np.cross(p[b] - p[a], p[c] - p[a])
        if np.dot(n, h.equations[k, :3]) < 0:
            tris[k] = [a, c, b]
    return tris.astype(np.uint32)


def from_convex_mesh(points, indices):
    """Returns dict(face_normal (nf,3) f32, face_first (nf,) u32, face_count (nf,) u32, vertices_adj_to_face (m,) u32,
    edges_adj_to_face (m,) u32, num_edges) or None where the reference returns None."""
    pts = np.asarray(points, dtype=np.float32)
    indices = np.asarray(indices, dtype=np.uint32)
    eps = F(np.sqrt(EPS))
    if len(pts) + len(indices) <= 2:
        return None
    edges = []       # dict(vertices, faces [t0, t1], deleted)
    edge_map = {}
    triangles = []   # dict(vertices, edges, normal, parent_face)
    for idx in indices:
        a, b, c = int(idx[0]), int(idx[1]), int(idx[2])
        if a == b or a == c or b == c:
            return None
        face_id = len(triangles)
        edges_id = [0xFFFFFFFF] * 3
        vs = [a, b, c]
        for i1 in range(3):
            i2 = (i1 + 1) % 3
            key = (min(vs[i1], vs[i2]), max(vs[i1], vs[i2]))
            if key in edge_map:
                e = edges[edge_map[key]]
                if e["faces"][1] == 0xFFFFFFFF:
                    edges_id[i1] = edge_map[key]
                    e["faces"][1] = face_id
                else:
                    return None   # t-junction
            else:
                edge_map[key] = len(edges)
                edges_id[i1] = len(edges)
                d = (pts[vs[i2]] - pts[vs[i1]]).astype(np.float32)
                _, n = _unit(d)
                edges.append({"vertices": (vs[i1], vs[i2]), "faces": [face_id, 0xFFFFFFFF], "deleted": bool(not (n > EPS))})
        ab = (pts[b] - pts[a]).astype(np.float32)
        ac = (pts[c] - pts[a]).astype(np.float32)
        cr = np.array([F(F(ab[1] * ac[2]) - F(ab[2] * ac[1])), F(F(ab[2] * ac[0]) - F(ab[0] * ac[2])), F(F(ab[0] * ac[1]) - F(ab[1] * ac[0]))], np.float32)
        nu, nn = _unit(cr)
        ok = bool(nn > EPS)
        triangles.append({"vertices": vs, "edges": edges_id, "normal": nu if ok else np.zeros(3, np.float32), "parent_face": None})
    for e in edges:
        if e["faces"][1] == 0xFFFFFFFF:
            return None
        n1, n2 = triangles[e["faces"][0]]["normal"], triangles[e["faces"][1]]["normal"]
        dot = F(F(F(n1[0] * n2[0]) + F(n1[1] * n2[1])) + F(n1[2] * n2[2]))
        if dot > F(1.0) - eps:
            e["deleted"] = True
    faces, edges_adj_to_face, vertices_adj_to_face = [], [], []
    for i in range(len(triangles)):
        if triangles[i]["parent_face"] is not None:
            continue
        for j1 in range(3):
            if edges[triangles[i]["edges"][j1]]["deleted"]:
                continue
            new_face_id = len(faces)
            first = len(edges_adj_to_face)
            count = 1
            edges_adj_to_face.append(triangles[i]["edges"][j1])
            vertices_adj_to_face.append(triangles[i]["vertices"][j1])
            start_vertex = triangles[i]["vertices"][j1]
            curr_triangle, curr_edge_id = i, (j1 + 1) % 3
            while triangles[curr_triangle]["vertices"][curr_edge_id] != start_vertex:
                curr_edge = triangles[curr_triangle]["edges"][curr_edge_id]
                curr_vertex = triangles[curr_triangle]["vertices"][curr_edge_id]
                triangles[curr_triangle]["parent_face"] = new_face_id
                if not edges[curr_edge]["deleted"]:
                    edges_adj_to_face.append(curr_edge)
                    vertices_adj_to_face.append(curr_vertex)
                    count += 1
                    curr_edge_id = (curr_edge_id + 1) % 3
                else:
                    f0, f1 = edges[curr_edge]["faces"]
                    curr_triangle = f1 if curr_triangle == f0 else f0
                    curr_edge_id = (triangles[curr_triangle]["edges"].index(curr_edge) + 1) % 3
                    assert triangles[curr_triangle]["vertices"][curr_edge_id] == curr_vertex
            if count > 2:
                faces.append({"first": first, "count": count, "normal": triangles[i]["normal"]})
            break
    if not faces:
        return None
    # vertices: adjacent faces / edges (convex_polyhedron.rs:569-617)
    nv = len(pts)
    vcount = np.zeros(nv, np.uint32)
    for f in faces:
        for v in vertices_adj_to_face[f["first"]:f["first"] + f["count"]]:
            vcount[v] += 1
    vfirst = np.concatenate([[0], np.cumsum(vcount)[:-1]]).astype(np.uint32)
    total = int(vcount.sum())
    faces_adj_to_vertex, edges_adj_to_vertex = np.zeros(total, np.uint32), np.zeros(total, np.uint32)
    fill = np.zeros(nv, np.uint32)
    for fid, f in enumerate(faces):
        for k in range(f["first"], f["first"] + f["count"]):
            v = vertices_adj_to_face[k]
            faces_adj_to_vertex[vfirst[v] + fill[v]] = fid
            edges_adj_to_vertex[vfirst[v] + fill[v]] = edges_adj_to_face[k]
            fill[v] += 1
    edge_dir = np.zeros((len(edges), 3), np.float32)
    for k, e in enumerate(edges):
        d = (pts[e["vertices"][1]] - pts[e["vertices"][0]]).astype(np.float32)
        u, n = _unit(d)
        edge_dir[k] = u if n > EPS else [1.0, 0.0, 0.0]   # Unit::try_new(.., DEFAULT_EPSILON).unwrap_or(x_axis)
    return {"vert_first": vfirst, "vert_count": vcount, "faces_adj_to_vertex": faces_adj_to_vertex, "edges_adj_to_vertex": edges_adj_to_vertex,
            "edge_dir": edge_dir,
            "face_normal": np.stack([f["normal"] for f in faces]).astype(np.float32),
            "face_first": np.array([f["first"] for f in faces], np.uint32), "face_count": np.array([f["count"] for f in faces], np.uint32),
            "vertices_adj_to_face": np.array(vertices_adj_to_face, np.uint32), "edges_adj_to_face": np.array(edges_adj_to_face, np.uint32),
            "num_edges": len(edges)}


def hull_table(hulls):
    """Concatenated topology of several hulls (list of (n_i, 3) point arrays), vertex / face / edge ids local to each hull:
    dict(hull_face_first, hull_face_count, face_normal, face_first, face_count, vertices_adj_to_face, edges_adj_to_face) plus the
    vertex side: vert_first / vert_count (one entry per point, hull after hull), faces_adj_to_vertex, edges_adj_to_vertex,
    hull_edge_first (per hull), edge_dir."""
    hf, hc, fn, ff, fc, va, ea = [], [], [], [], [], [], []
    vf, vc, fav, eav, ed, hef = [], [], [], [], [], []
    for p in hulls:
        t = from_convex_mesh(p, hull_triangles(p))
        assert t is not None
        hf.append(len(ff))
        hc.append(len(t["face_first"]))
        base = len(va)
        vf += [int(x) + len(fav) for x in t["vert_first"]]
        vc += [int(x) for x in t["vert_count"]]
        fav += [int(x) for x in t["faces_adj_to_vertex"]]
        eav += [int(x) for x in t["edges_adj_to_vertex"]]
        hef.append(sum(len(e) for e in ed))
        ed.append(t["edge_dir"])
        fn.append(t["face_normal"])
        ff += [int(x) + base for x in t["face_first"]]
        fc += [int(x) for x in t["face_count"]]
        va += [int(x) for x in t["vertices_adj_to_face"]]
        ea += [int(x) for x in t["edges_adj_to_face"]]
    return {"hull_face_first": np.array(hf, np.uint32), "hull_face_count": np.array(hc, np.uint32),
            "face_normal": np.concatenate(fn).astype(np.float32), "face_first": np.array(ff, np.uint32), "face_count": np.array(fc, np.uint32),
            "vertices_adj_to_face": np.array(va, np.uint32), "edges_adj_to_face": np.array(ea, np.uint32),
            # vertex side (support_feature_id_toward): per point of the concatenated hull points, in hull order
            "vert_first": np.array(vf, np.uint32), "vert_count": np.array(vc, np.uint32), "faces_adj_to_vertex": np.array(fav, np.uint32),
            "edges_adj_to_vertex": np.array(eav, np.uint32), "hull_edge_first": np.array(hef, np.uint32), "edge_dir": np.concatenate(ed).astype(np.float32)}
