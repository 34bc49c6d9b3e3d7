"""ctypes binding of include/rvh.h.  Fails loudly when librvh.so is missing."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

GRID_ON, WIND_A, WIND_B, GRID_INT32_WRAP, KEEP_CORRECTION, KEEP_ORDER = 1, 2, 4, 8, 16, 32
SDF_ON, REPULSION_ON, SDF_TMA = 64, 128, 256      # north-star extensions, not in the reference (include/rvh.h)

EXPORTED_SYMBOLS = [
    "rvh_default_config", "rvh_create", "rvh_nccl_unique_id", "rvh_create_sharded", "rvh_exchange_mode", "rvh_set_colliders",
    "rvh_upload_strands_aos", "rvh_init_synthetic_head", "rvh_import_strands_fd", "rvh_step", "rvh_step_n", "rvh_step_host",
    "rvh_download_strands_aos", "rvh_download_grid", "rvh_draw_indirect", "rvh_step_phases",
    "rvh_profile_enable", "rvh_profile_read", "rvh_sync", "rvh_last_step_ms", "rvh_kernel_launches",
    "rvh_last_error", "rvh_destroy", "rvh_collider_build", "rvh_collider_translate", "rvh_wind_fbm",
    "rvh_abi_version", "rvh_set_head_sdf", "rvh_bake_head_sdf_from_colliders", "rvh_bake_head_sdf_from_mesh",
    "rvh_download_head_sdf", "rvh_sdf_mode", "rvh_debug_hit_masks", "rvh_import_indirect_fd", "rvh_import_semaphore_fd",
    "rvh_debug_set_interop_device_buffers", "rvh_expand_strands", "rvh_expand_device_buffers", "rvh_init_from_mesh", "rvh_download_collider_mask",
]


class RvhError(RuntimeError):
    """The reference throws std::runtime_error on any failure (e.g. Renderer.cpp:2317-2319)."""


class RvhConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("num_strands", C.c_int), ("num_points", C.c_int), ("rest_length", C.c_float),
        ("gravity_y", C.c_float), ("damping", C.c_float), ("vmax", C.c_float), ("penalty_k", C.c_float),
        ("sphere_radius", C.c_float), ("grid_dim", C.c_int), ("grid_extent", C.c_float),
        ("grid_origin", C.c_float * 3), ("grid_scale", C.c_float), ("friction", C.c_float), ("flags", C.c_int),
        ("strands_per_thread", C.c_int), ("repulsion", C.c_float),
    ]


def library_path():
    return os.path.join(HERE, os.environ.get("RVH_LIB", "librvh.so"))      # RVH_LIB: tuning experiments only


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RvhError("librvh.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                       "there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    L.rvh_default_config.argtypes = [C.POINTER(RvhConfig), C.c_int, C.c_int]
    L.rvh_default_config.restype = None
    L.rvh_create.argtypes = [C.POINTER(vp), C.POINTER(RvhConfig)]
    L.rvh_nccl_unique_id.argtypes = [vp]
    L.rvh_create_sharded.argtypes = [C.POINTER(vp), C.POINTER(RvhConfig), C.c_int, C.c_int, vp]
    L.rvh_exchange_mode.argtypes = [vp]
    L.rvh_set_colliders.argtypes = [vp, vp, C.c_int]
    L.rvh_upload_strands_aos.argtypes = [vp, vp, C.c_size_t]
    L.rvh_init_synthetic_head.argtypes = [vp, C.c_ulonglong, C.c_float, C.c_ulonglong]
    L.rvh_import_strands_fd.argtypes = [vp, C.c_int, C.c_size_t]
    L.rvh_import_indirect_fd.argtypes = [vp, C.c_int, C.c_size_t]
    L.rvh_import_semaphore_fd.argtypes = [vp, C.c_int]
    L.rvh_debug_set_interop_device_buffers.argtypes = [vp, vp, C.c_size_t, vp]
    L.rvh_step.argtypes = [vp, C.c_float, C.c_float]
    L.rvh_step_n.argtypes = [vp, C.c_int, C.c_float, C.c_float, fp]
    L.rvh_step_host.argtypes = [vp, vp, C.c_size_t, C.c_float, C.c_float]
    L.rvh_download_strands_aos.argtypes = [vp, vp, C.c_size_t]
    L.rvh_download_grid.argtypes = [vp, vp, C.c_size_t]
    L.rvh_draw_indirect.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.rvh_step_phases.argtypes = [vp, C.c_float, C.c_float, C.c_int]
    L.rvh_profile_enable.argtypes = [vp, C.c_int]
    L.rvh_profile_read.argtypes = [vp, fp, C.POINTER(C.c_int)]
    L.rvh_sync.argtypes = [vp]
    L.rvh_last_step_ms.argtypes = [vp]
    L.rvh_last_step_ms.restype = C.c_float
    L.rvh_kernel_launches.argtypes = [vp]
    L.rvh_kernel_launches.restype = C.c_longlong
    L.rvh_last_error.argtypes = [vp]
    L.rvh_last_error.restype = C.c_char_p
    L.rvh_destroy.argtypes = [vp]
    L.rvh_destroy.restype = None
    L.rvh_collider_build.argtypes = [fp, fp, fp, fp]
    L.rvh_collider_build.restype = None
    L.rvh_collider_translate.argtypes = [fp, fp]
    L.rvh_collider_translate.restype = None
    L.rvh_wind_fbm.argtypes = [C.c_float]
    L.rvh_wind_fbm.restype = C.c_float
    ip = C.POINTER(C.c_int)
    L.rvh_set_head_sdf.argtypes = [vp, fp, ip, fp, C.c_float]
    L.rvh_bake_head_sdf_from_colliders.argtypes = [vp, ip, fp, C.c_float]
    L.rvh_bake_head_sdf_from_mesh.argtypes = [vp, fp, C.c_int, ip, C.c_int, ip, fp, C.c_float]
    L.rvh_download_head_sdf.argtypes = [vp, fp, C.c_size_t]
    L.rvh_sdf_mode.argtypes = [vp]
    L.rvh_debug_hit_masks.argtypes = [vp, C.POINTER(C.c_ubyte), C.c_size_t]
    L.rvh_download_collider_mask.argtypes = [vp, C.POINTER(C.c_ubyte), C.c_size_t, ip]
    L.rvh_expand_strands.argtypes = [vp, C.c_int, C.c_int, fp, fp, C.c_size_t, fp]
    L.rvh_init_from_mesh.argtypes = [vp, fp, fp, C.c_int, C.c_ulonglong, C.c_float, C.c_ulonglong]
    L.rvh_expand_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    _lib = L
    return L


def default_config(num_strands, num_points, flags=GRID_ON, device=0, rest_length=None, strands_per_thread=0, repulsion=None):
    cfg = RvhConfig()
    load_library().rvh_default_config(C.byref(cfg), num_strands, num_points)
    if repulsion is not None:
        cfg.repulsion = repulsion
    cfg.flags = flags
    cfg.device = device
    cfg.strands_per_thread = strands_per_thread
    if rest_length is not None:
        cfg.rest_length = rest_length
    return cfg


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def collider_build(trans, rot_deg, scale):
    t, r, s = (np.asarray(v, np.float32).copy() for v in (trans, rot_deg, scale))
    out = np.zeros(48, np.float32)
    load_library().rvh_collider_build(_fptr(t), _fptr(r), _fptr(s), _fptr(out))
    return out


def collider_translate(c48, translation):
    c = np.ascontiguousarray(c48, np.float32).copy()
    t = np.asarray(translation, np.float32).copy()
    load_library().rvh_collider_translate(_fptr(c), _fptr(t))
    return c


def wind_fbm(total_time):
    return float(load_library().rvh_wind_fbm(C.c_float(total_time)))


def nccl_unique_id():
    buf = (C.c_char * 128)()
    r = load_library().rvh_nccl_unique_id(C.cast(buf, C.c_void_p))
    if r != 0:
        raise RvhError("rvh_nccl_unique_id: %s" % load_library().rvh_last_error(None).decode())
    return bytes(buf)


class HairSim:
    """One context = one GPU's shard of strands (reference: Hair + Scene + the compute half of Renderer)."""

    def __init__(self, cfg, rank=0, nranks=1, nccl_id=None):
        self.L = load_library()
        self.cfg = cfg
        self.S, self.N = cfg.num_strands, cfg.num_points
        self.ctx = C.c_void_p()
        if nranks > 1:
            idbuf = C.create_string_buffer(nccl_id, 128)
            r = self.L.rvh_create_sharded(C.byref(self.ctx), C.byref(cfg), rank, nranks, C.cast(idbuf, C.c_void_p))
        else:
            r = self.L.rvh_create(C.byref(self.ctx), C.byref(cfg))
        if r != 0:
            raise RvhError("rvh_create failed (%d): %s" % (r, self.L.rvh_last_error(None).decode()))

    def _check(self, r, what):
        if r != 0:
            raise RvhError("%s failed (%d): %s" % (what, r, self.L.rvh_last_error(self.ctx).decode()))

    def close(self):
        if self.ctx:
            self.L.rvh_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def aos_bytes(self):
        return self.S * 48 * self.N

    def set_colliders(self, colliders):
        c = np.ascontiguousarray(colliders, np.float32).reshape(-1, 48)
        self._check(self.L.rvh_set_colliders(self.ctx, c.ctypes.data_as(C.c_void_p), c.shape[0]), "rvh_set_colliders")

    def upload(self, strands):
        a = np.ascontiguousarray(strands, np.float32)
        self._check(self.L.rvh_upload_strands_aos(self.ctx, a.ctypes.data_as(C.c_void_p), a.nbytes), "rvh_upload_strands_aos")

    def init_synthetic_head(self, first_strand=0, strand_length=2.5, seed=8):
        self._check(self.L.rvh_init_synthetic_head(self.ctx, first_strand, strand_length, seed), "rvh_init_synthetic_head")

    def init_from_mesh(self, tri_pos, tri_nrm=None, first_strand=0, strand_length=2.5, seed=8):
        """tri_pos / tri_nrm: float32 [ntris, 3, 3] corner positions / normals (normals optional)."""
        tp = np.ascontiguousarray(tri_pos, np.float32).reshape(-1, 9)
        tn = None if tri_nrm is None else np.ascontiguousarray(tri_nrm, np.float32).reshape(-1, 9)
        self._check(self.L.rvh_init_from_mesh(self.ctx, _fptr(tp), None if tn is None else _fptr(tn), tp.shape[0], first_strand, strand_length, seed),
                    "rvh_init_from_mesh")

    def upload_ptr(self, ptr, nbytes):
        self._check(self.L.rvh_upload_strands_aos(self.ctx, C.c_void_p(ptr), nbytes), "rvh_upload_strands_aos")

    def download_ptr(self, ptr, nbytes):
        self._check(self.L.rvh_download_strands_aos(self.ctx, C.c_void_p(ptr), nbytes), "rvh_download_strands_aos")

    def step_host_ptr(self, ptr, nbytes, dt, total_time):
        self._check(self.L.rvh_step_host(self.ctx, C.c_void_p(ptr), nbytes, dt, total_time), "rvh_step_host")

    def step_host(self, strands, dt, total_time=0.0):
        """In place on a C-contiguous float32 [S,3,N,4] array: upload, one step, download (rvh_step_host)."""
        assert strands.dtype == np.float32 and strands.flags["C_CONTIGUOUS"] and strands.nbytes == self.aos_bytes
        self.step_host_ptr(strands.ctypes.data, strands.nbytes, dt, total_time)
        return strands

    def step(self, dt, total_time=0.0):
        self._check(self.L.rvh_step(self.ctx, dt, total_time), "rvh_step")

    def step_phases(self, dt, total_time, phases):
        self._check(self.L.rvh_step_phases(self.ctx, dt, total_time, phases), "rvh_step_phases")

    def step_n(self, n, dt, total_time0=0.0, timed=True):
        ms = C.c_float(0.0)
        self._check(self.L.rvh_step_n(self.ctx, n, dt, total_time0, C.byref(ms) if timed else None), "rvh_step_n")
        return float(ms.value)

    def download(self):
        out = np.empty((self.S, 3, self.N, 4), np.float32)
        self._check(self.L.rvh_download_strands_aos(self.ctx, out.ctypes.data_as(C.c_void_p), out.nbytes), "rvh_download_strands_aos")
        return out

    def download_grid(self):
        G = self.cfg.grid_dim
        dt = np.int32 if (self.cfg.flags & GRID_INT32_WRAP) else np.int64
        out = np.empty((G ** 3, 4), dt)
        self._check(self.L.rvh_download_grid(self.ctx, out.ctypes.data_as(C.c_void_p), out.nbytes), "rvh_download_grid")
        return out

    def draw_indirect(self):
        out = (C.c_uint32 * 4)()
        self._check(self.L.rvh_draw_indirect(self.ctx, out), "rvh_draw_indirect")
        return list(out)

    def profile_enable(self, on=True):
        """False/0 off, True/1 every kernel, 2 only k_ftl_step."""
        self._check(self.L.rvh_profile_enable(self.ctx, int(on)), "rvh_profile_enable")

    def profile_read(self):
        ms = (C.c_float * 6)()
        n = (C.c_int * 6)()
        self._check(self.L.rvh_profile_read(self.ctx, ms, n), "rvh_profile_read")
        names = ["ftl_step", "grid_gather", "grid_allreduce", "grid_clear", "grid_splat", "grid_finalize"]
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(names)}

    # ---- head SDF (extension) ----
    @staticmethod
    def _dim_origin(dim, origin):
        d = np.asarray(dim, np.int32).copy()
        o = np.asarray(origin, np.float32).copy()
        return d, o, d.ctypes.data_as(C.POINTER(C.c_int)), _fptr(o)

    def set_head_sdf(self, vol, origin, cell):
        """vol: float32 [nz][ny][nx] signed distances at the lattice nodes (negative inside)."""
        v = np.ascontiguousarray(vol, np.float32)
        d, o, dp, op = self._dim_origin([v.shape[2], v.shape[1], v.shape[0]], origin)
        self._check(self.L.rvh_set_head_sdf(self.ctx, _fptr(v), dp, op, cell), "rvh_set_head_sdf")
        self._sdf_dim = tuple(int(x) for x in d)

    def bake_head_sdf_from_colliders(self, dim, origin, cell):
        d, o, dp, op = self._dim_origin(dim, origin)
        self._check(self.L.rvh_bake_head_sdf_from_colliders(self.ctx, dp, op, cell), "rvh_bake_head_sdf_from_colliders")
        self._sdf_dim = tuple(int(x) for x in d)

    def bake_head_sdf_from_mesh(self, verts, tris, dim, origin, cell):
        v = np.ascontiguousarray(verts, np.float32)
        t = np.ascontiguousarray(tris, np.int32)
        d, o, dp, op = self._dim_origin(dim, origin)
        self._check(self.L.rvh_bake_head_sdf_from_mesh(self.ctx, _fptr(v), v.shape[0], t.ctypes.data_as(C.POINTER(C.c_int)), t.shape[0], dp, op, cell),
                    "rvh_bake_head_sdf_from_mesh")
        self._sdf_dim = tuple(int(x) for x in d)

    def download_head_sdf(self):
        nx, ny, nz = self._sdf_dim
        out = np.empty((nz, ny, nx), np.float32)
        self._check(self.L.rvh_download_head_sdf(self.ctx, _fptr(out), out.nbytes), "rvh_download_head_sdf")
        return out

    def collider_mask(self):
        """uint8 [D, D, D] (z, y, x) candidate mask, or None when no mask is in use."""
        D = self.cfg.grid_dim // 2
        out = np.zeros((D, D, D), np.uint8)
        dim = C.c_int(0)
        self._check(self.L.rvh_download_collider_mask(self.ctx, out.ctypes.data_as(C.POINTER(C.c_ubyte)), out.nbytes, C.byref(dim)), "rvh_download_collider_mask")
        return out if dim.value else None

    def sdf_mode(self):
        return {0: "off", 1: "ldg", 2: "tma"}[int(self.L.rvh_sdf_mode(self.ctx))]

    def hit_masks(self):
        """uint8 [S, N]: the collider decisions this path takes for the state it holds (rvh_debug_hit_masks)."""
        out = np.zeros((self.S, self.N), np.uint8)
        self._check(self.L.rvh_debug_hit_masks(self.ctx, out.ctypes.data_as(C.POINTER(C.c_ubyte)), out.nbytes), "rvh_debug_hit_masks")
        return out

    def expand(self, isolines=12, divisions=42, download=True):
        """Guide -> render strands (hair.tesc/hair.tese).  Returns (pos_width, tangent_u, ms); arrays [S, isolines, divisions+1, 4]."""
        ms = C.c_float(0.0)
        if not download:
            self._check(self.L.rvh_expand_strands(self.ctx, isolines, divisions, None, None, 0, C.byref(ms)), "rvh_expand_strands")
            return None, None, float(ms.value)
        pw = np.empty((self.S, isolines, divisions + 1, 4), np.float32)
        tu = np.empty_like(pw)
        self._check(self.L.rvh_expand_strands(self.ctx, isolines, divisions, _fptr(pw), _fptr(tu), pw.nbytes, C.byref(ms)), "rvh_expand_strands")
        return pw, tu, float(ms.value)

    def exchange_mode(self):
        return {0: "single", 1: "nccl-allreduce", 2: "peer-memory-fused"}[int(self.L.rvh_exchange_mode(self.ctx))]

    def sync(self):
        self._check(self.L.rvh_sync(self.ctx), "rvh_sync")

    def kernel_launches(self):
        return int(self.L.rvh_kernel_launches(self.ctx))
