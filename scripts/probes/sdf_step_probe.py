"""One SDF step in TMA mode, small, for compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvh_b200 as rvh, orc
S, N, L = int(os.environ.get("S", 1024)), 8, 2.5
cols = rvh.scenes.bench_colliders()
rest = float(np.float32(L) / np.float32(N - 1))
vol = orc.sdf_bake_colliders(cols, [41, 63, 35], np.array([-2.0, -2.2, -1.8], np.float32), np.float32(0.1))
cfg = rvh.default_config(S, N, flags=rvh.SDF_ON | int(os.environ.get("XF", "0")), rest_length=rest)
sim = rvh.HairSim(cfg); sim.set_colliders(cols); sim.set_head_sdf(vol, [-2.0, -2.2, -1.8], 0.1)
print("mode", sim.sdf_mode(), flush=True)
sim.upload(rvh.scenes.synthetic_head(S, N, L)); sim.step(1 / 60, 0.0); out = sim.download(); print("ok", np.isfinite(out).all())
