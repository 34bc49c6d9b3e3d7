"""ctypes access to librvh_host.so (the C++ Hair/Scene/Renderer mirror, realtime-vulkan-hair_b200/host/)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "realtime-vulkan-hair_b200", "librvh_host.so")
        if not os.path.exists(path):
            raise RuntimeError("librvh_host.so not built: run __graft_entry__.build()")
        L = C.CDLL(path)
        fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.rvhh_last_error.restype = C.c_char_p
        L.rvhh_sizeof.argtypes = [C.c_int]
        L.rvhh_hair_init.argtypes = [C.c_char_p, C.c_int, C.c_int, fp, C.c_size_t, u32p]
        L.rvhh_run_scene.argtypes = [C.c_char_p, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp, C.c_size_t, u32p, fp]
        _lib = L
    return _lib


def write_obj(path, mesh):
    """OBJ text from the frozen mesh arrays (tests/golden/mannequin_segment_mesh.npz); %.9g round-trips float32."""
    with open(path, "w") as f:
        for x, y, z in mesh["v"]:
            f.write("v %.9g %.9g %.9g\n" % (x, y, z))
        for x, y, z in mesh["vn"]:
            f.write("vn %.9g %.9g %.9g\n" % (x, y, z))
        for a, b in zip(mesh["fv"], mesh["fn"]):
            f.write("f " + " ".join("%d//%d" % (i + 1, j + 1) for i, j in zip(a, b) if i >= 0) + "\n")


def hair_init(obj_path, S, N):
    st = np.zeros((S, 3, N, 4), np.float32)
    ind = (C.c_uint32 * 4)()
    n = lib().rvhh_hair_init(obj_path.encode(), S, N, st.ctypes.data_as(C.POINTER(C.c_float)), st.nbytes, ind)
    if n < 0:
        raise RuntimeError(lib().rvhh_last_error().decode())
    return st, list(ind)


def run_scene(obj_path, S, N, flags, frames, dt, sphere_moves=None, strands_in=None):
    out = np.zeros((S, 3, N, 4), np.float32)
    ind = (C.c_uint32 * 4)()
    tt = C.c_float(0)
    fp = C.POINTER(C.c_float)
    mv = None if sphere_moves is None else np.ascontiguousarray(sphere_moves, np.float32)
    si = None if strands_in is None else np.ascontiguousarray(strands_in, np.float32)
    r = lib().rvhh_run_scene((obj_path or "").encode(), None if si is None else si.ctypes.data_as(fp), S, N, flags, frames, dt,
                             None if mv is None else mv.ctypes.data_as(fp), out.ctypes.data_as(fp), out.nbytes, ind, C.byref(tt))
    if r != 0:
        raise RuntimeError(lib().rvhh_last_error().decode())
    return out, list(ind), float(tt.value)
