import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import parry_b200
from harness import scenes
FMAX = float(np.finfo(np.float32).max)
ctx = parry_b200.Context(0)
v, i = scenes.terrain(2001, 2001)
mesh = parry_b200.TriMesh(ctx, v, i)
m = 1 << 23
rays = torch.from_numpy(scenes.terrain_rays(m, seed=6)).pin_memory().numpy()
toi = torch.empty(m, dtype=torch.float32).pin_memory().numpy()
tri = torch.empty(m, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
for lg in (0, 18, 19):
    os.environ["PB2_RAY_CHUNK_LOG2"] = str(lg)  # 0 = built-in ramped schedule
    for _ in range(2): mesh.cast_local_ray(rays, FMAX, out=(toi, tri))
    t0 = time.perf_counter()
    for _ in range(5): mesh.cast_local_ray(rays, FMAX, out=(toi, tri))
    dt = (time.perf_counter() - t0) / 5
    print("chunk 2^%d: %.3f ms  %.1f Mrays/s" % (lg, dt * 1e3, m / dt / 1e6))
