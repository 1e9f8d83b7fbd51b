"""CPU check of device code that is written __host__ __device__: the core of the manifold-persistence kernel
(parry_b200/csrc/manifold_update.cuh) is compiled for the host by nvcc with the library's floating-point flags and must reproduce
the oracle's ContactManifold::try_update_contacts bit for bit. (The kernel around it is covered by tests/test_manifold_update_gpu.py.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from harness import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    from parry_b200 import build as b
    src = os.path.join(HERE, "hostcheck", "manifold_update_host.cu")
    out_dir = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    deps = [src, os.path.join(b.CSRC, "manifold_update.cuh"), os.path.join(b.CSRC, "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-shared", src, "-o", so])
    return C.CDLL(so)


def test_try_update_core_matches_oracle(hostlib, oracle):
    g = scenes.rng(51)
    T = oracle.ShapeTable([("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("ball", 0.3)])
    n = 20000
    s1, s2 = g.integers(0, 3, n).astype(np.uint32), g.integers(0, 3, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=8, threads=4)
    moved = p2.copy()
    moved[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.6, 3.0e-4, 0.02)).astype(np.float32)
    tilt = scenes.random_unit_quaternions(g, n)
    tilt[:, :3] *= 0.004                                   # a fraction of a degree on some pairs
    tilt /= np.linalg.norm(tilt, axis=1, keepdims=True)
    sel = g.random(n) < 0.3
    a, b = moved[sel, :4].astype(np.float64), tilt[sel]
    moved[sel, :4] = np.stack([a[:, 3] * b[:, 0] + a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                               a[:, 3] * b[:, 1] - a[:, 0] * b[:, 2] + a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0],
                               a[:, 3] * b[:, 2] + a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0] + a[:, 2] * b[:, 3],
                               a[:, 3] * b[:, 3] - a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2]], axis=1).astype(np.float32)
    kept_o, q_o = oracle.ShapeTable.manifolds_try_update(p1, moved, nr, cnt, pts)
    q_h = pts.copy()
    kept_h = np.zeros(n, np.uint8)
    P = C.c_void_p
    hostlib.hostcheck_manifolds_try_update.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_float, C.c_float, P, P, P, P]
    hostlib.hostcheck_manifolds_try_update(p1.ctypes.data, moved.ctypes.data, n, pts.shape[1], 0.99984769515, 1.0e-6, nr.ctypes.data,
                                           cnt.ctypes.data, q_h.ctypes.data, kept_h.ctypes.data)
    has = cnt > 0
    assert 0.2 < kept_o[has].mean() < 0.9                  # both outcomes exercised
    assert (kept_h == kept_o).all()
    assert (q_h.view(np.uint32) == q_o.view(np.uint32)).all()   # including the partially refreshed points of rejected manifolds
