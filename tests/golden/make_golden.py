#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's OWN sources compiled under oracle/_ref
(see oracle/ref_build/build_ref.py).  Run in the build container, where /root/reference
exists:   python oracle/ref_build/build_ref.py && python tests/golden/make_golden.py

c1_reference_scene.npz  (config C1: the shipped scene, S=900, N=10, dt=1/60, grid on/int32)
  colliders        [6,48]  reference Collider ctor (Scene.h:28-38) on main.cpp:229-237 values
  sphere_moved     [48]    collider 0 after translateSphere((0.5,0.25,-0.125)) (Scene.cpp:110-120)
  state0           [900,3,10,4]  what the reference's Hair::Hair uploads (Strand.cpp:149-191)
  indirect0        [4]
  k                [5]     step indices 0,1,10,100,299
  pre_k / post_k   [5,900,2,10,3] pos+vel xyz before / after ONE dispatch of the reference shader
                   text; state k is reached by free-running that shader from state0
  corr_post        [5,900,10,3]  correctionVecs after that dispatch
  grid_idx_k / grid_val_k   sparse int32 GridCell contents after the dispatch at step k
wind_n32.npz  (wind lines un-commented, N=32 parameterised build, 256 strands, 1 dispatch)
"""
import os
import sys
import ctypes as C

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

OBJ = b"/root/reference/src/models/mannequin_segment.obj"
TRS = [((2.0, 0.0, 1.0), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
       ((0.0, 2.64, 0.08), (-38.270, 0.0, 0.0), (0.817, 1.158, 1.01)),
       ((0.0, 1.35, -0.288), (18.301, 0.0, 0.0), (0.457, 1.0, 0.538)),
       ((0.0, -0.380, -0.116), (-17.260, 0.0, 0.0), (1.078, 1.683, 0.974)),
       ((-0.698, 0.087, -0.36), (-20.254, 13.144, 34.5), (0.721, 1.0, 0.724)),
       ((0.698, 0.087, -0.36), (-20.254, 13.144, -34.5), (0.721, 1.0, 0.724))]


def main():
    H = orc.ref_host()
    cols = np.zeros((6, 48), np.float32)
    for i, (t, r, s) in enumerate(TRS):
        t, r, s = (np.array(v, np.float32) for v in (t, r, s))
        H.ref_collider_build(orc._f(t), orc._f(r), orc._f(s), orc._f(cols[i]))
    moved = cols[0].copy()
    tr = np.array([0.5, 0.25, -0.125], np.float32)
    H.ref_collider_translate(orc._f(moved), orc._f(tr))

    st0 = np.zeros((900, 3, 10, 4), np.float32)
    ind0 = np.zeros(4, np.uint32)
    n = H.ref_hair_init(OBJ, orc._f(st0), st0.nbytes, ind0.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert n == 900

    dt = np.float32(1.0 / 60.0)
    ks = [0, 1, 10, 100, 299]
    pre, post, corr, gidx, gval = [], [], [], [], []
    st = st0.copy()
    for k in range(300):
        nxt, grid, ind = orc.ref_dispatch("N10", st, cols, dt, np.float32(k) * dt)
        if k in ks:
            pre.append(st[:, 0:2, :, :3].copy())
            post.append(nxt[:, 0:2, :, :3].copy())
            corr.append(nxt[:, 2, :, :3].copy())
            nz = np.flatnonzero(np.any(grid != 0, axis=1))
            gidx.append(nz.astype(np.int32))
            gval.append(grid[nz].copy())
            assert np.all(nxt[:, 0, :, 3] == 1.0) and np.all(nxt[:, 1, :, 3] == 0.0)
        st = nxt
    out = dict(colliders=cols, sphere_moved=moved, sphere_translation=tr, state0=st0, indirect0=ind0,
               k=np.array(ks, np.int32), pre=np.stack(pre), post=np.stack(post), corr_post=np.stack(corr), dt=dt)
    for i, k in enumerate(ks):
        out["grid_idx_%d" % k] = gidx[i]
        out["grid_val_%d" % k] = gval[i]
    np.savez_compressed(os.path.join(HERE, "c1_reference_scene.npz"), **out)

    # wind variants at N=32 (parameterised build of the same shader text)
    rng = np.random.default_rng(8)
    S, N = 256, 32
    roots = np.stack([rng.uniform(-0.6, 0.6, S), rng.uniform(3.2, 3.7, S), rng.uniform(-0.6, 0.6, S)], 1).astype(np.float32)
    dirs = rng.normal(size=(S, 3)).astype(np.float32)
    dirs[:, 1] = -np.abs(dirs[:, 1]) - 0.5
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rest = np.float32(2.5) / np.float32(N - 1)
    stw = np.zeros((S, 3, N, 4), np.float32)
    stw[:, 0, :, :3] = roots[:, None, :] + (np.arange(N, dtype=np.float32) * rest)[None, :, None] * dirs[:, None, :]
    stw[:, 0, :, 3] = 1.0
    stw[:, 1, :, :3] = rng.normal(scale=0.5, size=(S, N, 3)).astype(np.float32)
    wout = dict(state=stw, colliders=cols, dt=dt, total_time=np.float32(1.2345))
    for w in ("A", "B"):
        nxt, grid, _ = orc.ref_dispatch("N32_wind" + w, stw, cols, dt, np.float32(1.2345))
        wout["post_" + w] = nxt
    nxt, grid, _ = orc.ref_dispatch("N32", stw, cols, dt, np.float32(1.2345))
    wout["post_none"] = nxt
    np.savez_compressed(os.path.join(HERE, "wind_n32.npz"), **wout)
    for f in ("c1_reference_scene.npz", "wind_n32.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
