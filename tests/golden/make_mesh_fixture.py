#!/usr/bin/env python3
"""Freeze the follicle surface of the shipped scene (reference asset src/models/mannequin_segment.obj,
main.cpp:226) as arrays, so that the host mirror's Hair constructor can be pinned against the reference's
own Hair::Hair output (state0 in c1_reference_scene.npz) on boxes where /root/reference does not exist.
Run in the build container:  python tests/golden/make_mesh_fixture.py

mannequin_segment_mesh.npz
  v   [nv,3]  float32   vertex positions, file order
  vn  [nn,3]  float32   vertex normals, file order
  fv  [F,4]   int32     per-face vertex indices (0-based; -1 pads faces with fewer than 4 corners)
  fn  [F,4]   int32     per-face normal indices
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = "/root/reference/src/models/mannequin_segment.obj"
HEAD_OBJ = "/root/reference/src/models/mannequin.obj"      # the rendered mannequin (main.cpp:222-223): head-SDF bake fixture


def head_mesh():
    """mannequin_head_mesh.npz: v [nv,3] float32, tri [nt,3] int32 (quads split 0-1-2 / 0-2-3), positions only."""
    v, tri = [], []
    for ln in open(HEAD_OBJ):
        t = ln.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            a = [int(tok.split("/")[0]) - 1 for tok in t[1:]]
            for k in range(1, len(a) - 1):
                tri.append([a[0], a[k], a[k + 1]])
    out = os.path.join(HERE, "mannequin_head_mesh.npz")
    np.savez_compressed(out, v=np.array(v, np.float32), tri=np.array(tri, np.int32))
    print(out, os.path.getsize(out), len(v), len(tri))


def main():
    head_mesh()
    v, vn, fv, fn = [], [], [], []
    for ln in open(OBJ):
        t = ln.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "vn":
            vn.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            a, b = [], []
            for tok in t[1:]:
                p = tok.split("/")
                a.append(int(p[0]) - 1)
                b.append(int(p[2]) - 1)
            assert len(a) <= 4
            fv.append(a + [-1] * (4 - len(a)))
            fn.append(b + [-1] * (4 - len(b)))
    out = os.path.join(HERE, "mannequin_segment_mesh.npz")
    np.savez_compressed(out, v=np.array(v, np.float32), vn=np.array(vn, np.float32), fv=np.array(fv, np.int32), fn=np.array(fn, np.int32))
    print(out, os.path.getsize(out), len(v), len(vn), len(fv))


if __name__ == "__main__":
    main()
