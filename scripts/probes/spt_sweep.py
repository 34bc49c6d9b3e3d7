"""Strands per thread (1 = scalar, 2 = fp32x2 packs) over the strand count: ms per resident step, grid + wind B + colliders."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rvh_b200 as rvh

DT = float(np.float32(1.0 / 60.0))
cols = rvh.scenes.bench_colliders()
for S, N in [(100000, 64), (150000, 32), (250000, 32), (400000, 32), (600000, 32), (200000, 16), (400000, 16)]:
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    row = []
    for spt in (1, 2):
        sim = rvh.HairSim(rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B, strands_per_thread=spt))
        sim.set_colliders(cols); sim.upload(st)
        sim.step_n(10, DT, 0.0, timed=True)
        ms = min(sim.step_n(40, DT, 0.2, timed=True) / 40 for _ in range(3))
        sim.close()
        row.append(ms)
    print("S=%d N=%d  spt1 %.4f ms  spt2 %.4f ms  -> %s" % (S, N, row[0], row[1], "1" if row[0] < row[1] else "2"), flush=True)
