#!/usr/bin/env python3
"""Freeze outputs of the reference's OWN hair.tese text (oracle/_ref/libref_tese_N10.so, built by oracle/ref_build/build_ref.py
from /root/reference/src/shaders/hair.tese) as tests/golden/expand_tese_n10.npz: 24 guide strands x 12 isolines x 42 divisions
(j = 42, v = 1, is left out: the shader reads one curve point past the end there).  Run in the build container only."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvh_b200 as rvh  # noqa: E402


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_tese_N10.so"))
    fp = C.POINTER(C.c_float)
    L.ref_tese_eval.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]
    S, N, I, D = 24, 10, 12, 42
    st = rvh.scenes.synthetic_head(S, N, 2.5)
    rng = np.random.default_rng(2)
    st[:, 0, 1:, :3] += rng.normal(scale=0.02, size=(S, N - 1, 3)).astype(np.float32)       # bent guides
    ref = np.zeros((S, I, D, 8), np.float32)
    out = np.zeros(8, np.float32)
    for s in range(S):
        pts = np.ascontiguousarray(st[s, 0])
        for k in range(I):
            for j in range(D):
                L.ref_tese_eval(pts.ctypes.data_as(fp), I, D, k, j, out.ctypes.data_as(fp))
                ref[s, k, j] = out
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "expand_tese_n10.npz"), state=st[:, 0:2].copy(), ref=ref, isolines=I, divisions=D)
    print("wrote expand_tese_n10.npz", ref.shape)


if __name__ == "__main__":
    main()
