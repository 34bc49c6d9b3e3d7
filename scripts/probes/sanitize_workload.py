"""Small run through every kernel of librvh.so, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvh_b200 as rvh

DT = float(np.float32(1 / 60))
cols = rvh.scenes.bench_colliders()
origin, cell, dim = np.array([-2.0, -2.2, -1.8], np.float32), 0.1, [41, 63, 35]
m = np.load(os.path.join(ROOT, "tests", "golden", "mannequin_head_mesh.npz"))
for S, N, spt in ((2048, 12, 1), (1500, 10, 2), (2048, 6, 4)):
    rest = float(np.float32(2.5) / np.float32(N - 1))
    for flags in (rvh.GRID_ON | rvh.WIND_B, rvh.GRID_ON | rvh.SDF_ON | rvh.REPULSION_ON, rvh.GRID_ON | rvh.SDF_ON | rvh.SDF_TMA | rvh.WIND_A, rvh.KEEP_CORRECTION):
        sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=rest, strands_per_thread=spt))
        sim.set_colliders(cols)
        if flags & rvh.SDF_ON:
            if flags & rvh.SDF_TMA:
                sim.bake_head_sdf_from_colliders(dim, origin, cell)
            else:
                sim.bake_head_sdf_from_mesh(m["v"] * np.float32(0.98), m["tri"][:600], dim, origin, cell)
        sim.upload(rvh.scenes.synthetic_head(S, N, 2.5))
        for k in range(3):
            sim.step(DT, 0.1 * k)
        out = sim.download()
        sim.step_phases(DT, 0.0, 1); sim.step_phases(DT, 0.0, 2)
        g = sim.download_grid()
        pw, tu, ms = sim.expand(5, 9)
        sim.init_synthetic_head(0, 2.5, 8)
        sim.step(DT, 0.0)
        assert np.isfinite(sim.download()).all() and np.isfinite(pw).all()
        sim.close()
        print("ok", S, N, spt, flags, flush=True)
