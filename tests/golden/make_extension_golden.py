#!/usr/bin/env python3
"""Freeze outputs of the extension oracle modes (oracle.c ORC_SDF_ON / ORC_REPULSION_ON / orc_expand_strands) on a small
scene, so that later changes to the oracle or the kernels are caught against fixed numbers.  The reference has nothing to
pin these against (SURVEY.md top table): the vectors are regression pins of THIS repository's definitions, produced by
the oracle built with -ffp-contract=off.      python tests/golden/make_extension_golden.py

extensions.npz
  colliders [6,48], sdf_dim [3], sdf_origin [3], sdf_cell                   lattice of the volume baked from the ellipsoids
  sdf_slice [ny,nx], sdf_sum                                                its middle z slice and float64 sum (the full volume is re-baked by the test)
  pre [S,2,N,3]                 state after 25 oracle steps (hair resting on the head)
  post_sdf / post_rep / post_both [S,2,N,3]   one more step with SDF_ON|GRID_ON, GRID_ON|REPULSION_ON, all three
  exp_pw / exp_tu [S,4,10,4]    expansion of `pre` with 4 isolines x 9 divisions
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc                    # noqa: E402
import rvh_b200 as rvh        # noqa: E402

S, N, L = 96, 12, 2.5
DT = np.float32(1.0 / 60.0)


def main():
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    dim, origin, cell = [41, 63, 35], np.array([-2.0, -2.2, -1.8], np.float32), np.float32(0.1)
    vol = orc.sdf_bake_colliders(cols, dim, origin, cell)
    orc.set_head_sdf(vol, origin, cell)
    st = rvh.scenes.synthetic_head(S, N, L)
    p = orc.default_params(S, N, orc.SDF_ON | orc.GRID_ON, rest_length=rest)
    for k in range(25):
        st, _ = orc.step(p, cols, DT, 0.0, st)
    out = {}
    for name, fl in (("post_sdf", orc.SDF_ON | orc.GRID_ON), ("post_rep", orc.GRID_ON | orc.REPULSION_ON), ("post_both", orc.SDF_ON | orc.GRID_ON | orc.REPULSION_ON)):
        q = orc.default_params(S, N, fl, rest_length=rest)
        o, _ = orc.step(q, cols, DT, 0.0, st)
        out[name] = o[:, 0:2, :, :3].copy()
    pw, tu = orc.expand_strands(st, 4, 9)
    path = os.path.join(HERE, "extensions.npz")
    np.savez_compressed(path, colliders=cols, sdf_dim=np.array(dim, np.int32), sdf_origin=origin, sdf_cell=cell, sdf_slice=vol[dim[2] // 2].copy(), sdf_sum=np.float64(vol.astype(np.float64).sum()),
                        pre=st[:, 0:2, :, :3].copy(), exp_pw=pw, exp_tu=tu, **out)
    inside = sum(1 for x in st[:, 0, 1:, :3].reshape(-1, 3) if (lambda r: r[0] and r[1] < 0)(orc.sdf_sample(x)))
    print(path, os.path.getsize(path), "points inside the SDF in `pre`:", inside)


if __name__ == "__main__":
    main()
