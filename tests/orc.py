"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built, the
reference-compiled checkers under oracle/_ref/.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

GRID_ON, WIND_A, WIND_B, GRID_INT32_WRAP = 1, 2, 4, 8
SDF_ON, REPULSION_ON = 64, 128          # extensions (not in the reference), see oracle.h


class OrcParams(C.Structure):
    _fields_ = [
        ("num_strands", C.c_int), ("num_points", C.c_int), ("rest_length", C.c_float),
        ("gravity_y", C.c_float), ("damping", C.c_float), ("vmax", C.c_float),
        ("penalty_k", C.c_float), ("sphere_radius", C.c_float), ("grid_dim", C.c_int),
        ("grid_extent", C.c_float), ("grid_origin", C.c_float * 3), ("grid_scale", C.c_float),
        ("friction", C.c_float), ("flags", C.c_int), ("num_colliders", C.c_int), ("repulsion", C.c_float),
    ]


_fp = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_default_params.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int]
        L.orc_collider_build.argtypes = [_fp, _fp, _fp, _fp]
        L.orc_collider_translate.argtypes = [_fp, _fp]
        L.orc_default_colliders.argtypes = [_fp]
        L.orc_fbm_time.argtypes = [C.c_float]
        L.orc_fbm_time.restype = C.c_float
        L.orc_step.argtypes = [C.POINTER(OrcParams), _fp, C.c_float, C.c_float, _fp, _i64p]
        L.orc_phase_integrate.argtypes = [C.POINTER(OrcParams), _fp, C.c_float, C.c_float, _fp]
        L.orc_phase_splat.argtypes = [C.POINTER(OrcParams), C.c_float, _fp, _i64p]
        L.orc_phase_gather.argtypes = [C.POINTER(OrcParams), _fp, _i64p]
        L.orc_step_parallel.argtypes = [C.POINTER(OrcParams), _fp, C.c_float, C.c_float, _fp, _i64p, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_init_strands_reference.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp]
        _ip = C.POINTER(C.c_int)
        L.orc_set_head_sdf.argtypes = [_fp, _ip, _fp, C.c_float]
        L.orc_sdf_sample.argtypes = [_fp, _fp, _fp]
        L.orc_sdf_sample.restype = C.c_int
        L.orc_sdf_bake_colliders.argtypes = [_fp, C.c_int, _ip, _fp, C.c_float, _fp]
        L.orc_sdf_bake_mesh.argtypes = [_fp, _ip, C.c_int, _ip, _fp, C.c_float, _fp]
        L.orc_expand_strands.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp]
        L.orc_hit_masks.argtypes = [C.POINTER(OrcParams), _fp, _fp, C.POINTER(C.c_ubyte)]
        _lib = L
    return _lib


def default_params(S, N, flags=GRID_ON, rest_length=None):
    p = OrcParams()
    lib().orc_default_params(C.byref(p), S, N)
    p.flags = flags
    if rest_length is not None:
        p.rest_length = rest_length
    return p


def default_colliders():
    out = np.zeros((6, 48), np.float32)
    lib().orc_default_colliders(_f(out))
    return out


def collider_build(trans, rot, scale):
    out = np.zeros(48, np.float32)
    t, r, s = (np.asarray(v, np.float32).copy() for v in (trans, rot, scale))
    lib().orc_collider_build(_f(t), _f(r), _f(s), _f(out))
    return out


def collider_translate(c48, translation):
    c = np.ascontiguousarray(c48, np.float32).copy()
    t = np.asarray(translation, np.float32).copy()
    lib().orc_collider_translate(_f(c), _f(t))
    return c


def fbm_time(t):
    return float(lib().orc_fbm_time(C.c_float(t)))


def new_grid(p):
    return np.zeros((p.grid_dim ** 3, 4), np.int64)


def step(p, colliders, dt, total_time, strands, grid=None, threads=0):
    """One full step in place on a COPY; returns (strands, grid)."""
    st = np.ascontiguousarray(strands, np.float32).copy()
    g = new_grid(p) if grid is None else grid
    col = np.ascontiguousarray(colliders, np.float32)
    if threads:
        lib().orc_step_parallel(C.byref(p), _f(col), dt, total_time, _f(st), g.ctypes.data_as(_i64p), threads)
    else:
        lib().orc_step(C.byref(p), _f(col), dt, total_time, _f(st), g.ctypes.data_as(_i64p))
    return st, g


def phase_integrate(p, colliders, dt, total_time, strands):
    st = np.ascontiguousarray(strands, np.float32).copy()
    col = np.ascontiguousarray(colliders, np.float32)
    lib().orc_phase_integrate(C.byref(p), _f(col), dt, total_time, _f(st))
    return st


def phase_splat(p, dt, strands):
    st = np.ascontiguousarray(strands, np.float32).copy()
    g = new_grid(p)
    lib().orc_phase_splat(C.byref(p), dt, _f(st), g.ctypes.data_as(_i64p))
    return st, g


def phase_gather(p, strands, grid):
    st = np.ascontiguousarray(strands, np.float32).copy()
    g = np.ascontiguousarray(grid, np.int64)
    lib().orc_phase_gather(C.byref(p), _f(st), g.ctypes.data_as(_i64p))
    return st


def hit_masks(p, colliders, strands):
    """uint8 [S, N]: the oracle's collider decisions for `strands` (bit 0 sphere, bit j collider j)."""
    st = np.ascontiguousarray(strands, np.float32)
    col = np.ascontiguousarray(colliders, np.float32)
    out = np.zeros((st.shape[0], st.shape[2]), np.uint8)
    lib().orc_hit_masks(C.byref(p), _f(col), _f(st), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out


def init_strands_reference(roots, normals, N):
    S = roots.shape[0]
    st = np.zeros((S, 3, N, 4), np.float32)
    r = np.ascontiguousarray(roots, np.float32)
    n = np.ascontiguousarray(normals, np.float32)
    lib().orc_init_strands_reference(S, N, _f(r), _f(n), _f(st))
    return st


# ---- extensions: head SDF ----------------------------------------------------------------

_sdf_keep = None


def set_head_sdf(vol, origin, cell):
    """vol: float32 [nz][ny][nx] (None clears).  The oracle keeps a pointer, so the array is pinned here."""
    global _sdf_keep
    if vol is None:
        lib().orc_set_head_sdf(None, None, None, 0.0)
        _sdf_keep = None
        return
    v = np.ascontiguousarray(vol, np.float32)
    dim = np.array([v.shape[2], v.shape[1], v.shape[0]], np.int32)
    o = np.asarray(origin, np.float32).copy()
    _sdf_keep = (v, dim, o)
    lib().orc_set_head_sdf(_f(v), dim.ctypes.data_as(C.POINTER(C.c_int)), _f(o), C.c_float(cell))


def sdf_sample(p):
    pp = np.asarray(p, np.float32).copy()
    d = np.zeros(1, np.float32)
    g = np.zeros(3, np.float32)
    ok = lib().orc_sdf_sample(_f(pp), _f(d), _f(g))
    return bool(ok), float(d[0]), g


def sdf_bake_colliders(colliders, dim, origin, cell):
    col = np.ascontiguousarray(colliders, np.float32).reshape(-1, 48)
    d = np.asarray(dim, np.int32).copy()
    o = np.asarray(origin, np.float32).copy()
    out = np.zeros((d[2], d[1], d[0]), np.float32)
    lib().orc_sdf_bake_colliders(_f(col), col.shape[0], d.ctypes.data_as(C.POINTER(C.c_int)), _f(o), C.c_float(cell), _f(out))
    return out


def sdf_bake_mesh(verts, tris, dim, origin, cell):
    v = np.ascontiguousarray(verts, np.float32)
    t = np.ascontiguousarray(tris, np.int32)
    d = np.asarray(dim, np.int32).copy()
    o = np.asarray(origin, np.float32).copy()
    out = np.zeros((d[2], d[1], d[0]), np.float32)
    lib().orc_sdf_bake_mesh(_f(v), t.ctypes.data_as(C.POINTER(C.c_int)), t.shape[0], d.ctypes.data_as(C.POINTER(C.c_int)), _f(o), C.c_float(cell), _f(out))
    return out


def expand_strands(strands, isolines=12, divisions=42):
    """hair.tesc/hair.tese restated: (pos_width, tangent_u), each [S, isolines, divisions+1, 4]."""
    st = np.ascontiguousarray(strands, np.float32)
    S, _, N, _ = st.shape
    pw = np.zeros((S, isolines, divisions + 1, 4), np.float32)
    tu = np.zeros_like(pw)
    lib().orc_expand_strands(_f(st), S, N, isolines, divisions, _f(pw), _f(tu))
    return pw, tu


# ---- oracle/_ref: the reference's own sources compiled here ---------------------------

def ref_available(tag="N10"):
    return os.path.exists(os.path.join(REF_DIR, "libref_compute_%s.so" % tag))


def ref_host_available():
    return os.path.exists(os.path.join(REF_DIR, "libref_host.so"))


_ref_host = None
_ref_compute = {}


def ref_host():
    global _ref_host
    if _ref_host is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_host.so"))
        L.ref_hair_init.argtypes = [C.c_char_p, _fp, C.c_size_t, C.POINTER(C.c_uint32)]
        L.ref_collider_build.argtypes = [_fp, _fp, _fp, _fp]
        L.ref_collider_translate.argtypes = [_fp, _fp]
        _ref_host = L
    return _ref_host


def ref_compute(tag="N10"):
    if tag not in _ref_compute:
        L = C.CDLL(os.path.join(REF_DIR, "libref_compute_%s.so" % tag))
        L.ref_compute_dispatch.argtypes = [C.c_int, _fp, _fp, C.c_float, C.c_float, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_uint32), C.c_int]
        _ref_compute[tag] = L
    return _ref_compute[tag]


def ref_dispatch(tag, strands, colliders, dt, total_time, emulate_oob=0):
    """Run the reference shader text (compiled against glm) for one dispatch."""
    L = ref_compute(tag)
    st = np.ascontiguousarray(strands, np.float32).copy()
    S = st.shape[0]
    assert st.shape[2] == L.ref_shader_num_curve_points()
    G = L.ref_shader_grid_dim()
    grid = np.zeros((G ** 3, 4), np.int32)
    ind = np.zeros(4, np.uint32)
    col = np.ascontiguousarray(colliders, np.float32)
    n = L.ref_compute_dispatch(S, _f(st), _f(col), dt, total_time, grid.ctypes.data_as(C.POINTER(C.c_int32)),
                               ind.ctypes.data_as(C.POINTER(C.c_uint32)), emulate_oob)
    assert n >= S
    return st, grid, ind
