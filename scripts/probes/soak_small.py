"""Soak of the small-scene fast paths: many launches of k_scene_step (C1) and k_ftl_wave (C2); state finite, segment lengths kept,
and the persistent-kernel result still bit-identical to launch-per-kernel stepping at the end."""
import os, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rvh_b200 as rvh

DT = float(np.float32(1.0 / 60.0))
g = np.load("tests/golden/c1_reference_scene.npz")
st, cols = g["state0"], g["colliders"]
flags = rvh.GRID_ON | rvh.GRID_INT32_WRAP
a = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); a.set_colliders(cols); a.upload(st)
os.environ["RVH_SCENE_CTAS"] = "0"
b = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); b.set_colliders(cols); b.upload(st)
del os.environ["RVH_SCENE_CTAS"]
t0 = time.time(); l0 = a.kernel_launches()
total = 0
for rep in range(40):
    a.step_n(2500, DT, 0.0); total += 2500
a.sync()
print("k_scene_step: %d steps in %d launches, %.2f s" % (total, a.kernel_launches() - l0, time.time() - t0), flush=True)
for rep in range(40):
    b.step_n(2500, DT, 0.0)
fa, fb = a.download(), b.download()
assert np.isfinite(fa).all()
seg = np.linalg.norm(fa[:, 0, 1:, :3] - fa[:, 0, :-1, :3], axis=2)
print("segment length max rel dev %.2e" % np.abs(seg / np.float32(2.5 / 9) - 1).max())
print("bit-identical to launch-per-kernel after %d steps: %s; grids equal: %s" % (total, np.array_equal(fa.view(np.uint32), fb.view(np.uint32)), np.array_equal(a.download_grid(), b.download_grid())), flush=True)
a.close(); b.close()

S, N = 16384, 32
c = rvh.HairSim(rvh.default_config(S, N, flags=rvh.WIND_B)); c.set_colliders(rvh.scenes.bench_colliders()); c.upload(rvh.scenes.synthetic_head(S, N, 2.5))
t0 = time.time()
for rep in range(20):
    c.step_n(2500, DT, 0.01 * rep)
f = c.download()
print("k_ftl_wave: 50000 steps, %.2f s, finite: %s" % (time.time() - t0, bool(np.isfinite(f).all())), flush=True)
c.close()
