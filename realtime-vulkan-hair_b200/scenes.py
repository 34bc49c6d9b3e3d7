"""Scene inputs for tests and benchmarks (host side, numpy).

* `reference_colliders()`  -- the six colliders of the shipped scene (main.cpp:229-237), built
  with the library's own Collider maths (rvh_collider_build <-> Scene.h:28-38).
* `synthetic_head(S, N, L)` -- seeded synthetic head of SURVEY.md section 8(d): roots on the top
  hemisphere of the head ellipsoid (collider 1), counter-based splitmix64 (seed 8) keyed by the
  GLOBAL strand id so any rank can generate exactly its shard, points at exact rest spacing.
"""
import numpy as np

from .binding import collider_build, collider_translate

# main.cpp:229-237: translation, rotation (degrees), scale
REFERENCE_COLLIDER_TRS = [
    ((2.0, 0.0, 1.0), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),                       # sphere (movable)
    ((0.0, 2.64, 0.08), (-38.270, 0.0, 0.0), (0.817, 1.158, 1.01)),            # head
    ((0.0, 1.35, -0.288), (18.301, 0.0, 0.0), (0.457, 1.0, 0.538)),            # neck
    ((0.0, -0.380, -0.116), (-17.260, 0.0, 0.0), (1.078, 1.683, 0.974)),       # bust
    ((-0.698, 0.087, -0.36), (-20.254, 13.144, 34.5), (0.721, 1.0, 0.724)),    # right shoulder
    ((0.698, 0.087, -0.36), (-20.254, 13.144, -34.5), (0.721, 1.0, 0.724)),    # left shoulder
]


def reference_colliders():
    return np.stack([collider_build(t, r, s) for (t, r, s) in REFERENCE_COLLIDER_TRS]).astype(np.float32)


def bench_colliders():
    """Reference colliders with the sphere moved to (0.9, 2.2, 0.5) so that it intersects the hair."""
    c = reference_colliders()
    cur = c[0, 12:15].copy()
    c[0] = collider_translate(c[0], np.array([0.9, 2.2, 0.5], np.float32) - cur)
    return c


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _head_frames(colliders):
    head = np.asarray(colliders, np.float32).reshape(-1, 48)[1]
    return head[0:16].reshape(4, 4).T.astype(np.float32), head[32:48].reshape(4, 4).T.astype(np.float32)   # column-major -> [row][col]


def _head_roots(sid, X, seed):
    """Unit direction and root of the strands with global ids `sid` (uint64): the id is hashed, so consecutive ids are scattered."""
    base = (np.uint64(seed) << np.uint64(32)) + np.uint64(2) * sid
    r0 = (_splitmix64(base) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)
    r1 = (_splitmix64(base + np.uint64(1)) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)
    uy = r0
    rad = np.sqrt(np.maximum(np.float32(1.0) - uy * uy, np.float32(0.0))).astype(np.float32)
    phi = (np.float32(2.0 * np.pi) * r1).astype(np.float32)
    u = np.stack([rad * np.cos(phi), uy, rad * np.sin(phi)], axis=1).astype(np.float32)
    return u, (u @ X[:3, :3].T + X[:3, 3]).astype(np.float32)


def spatial_shard_ids(num_strands_total, rank, nranks, colliders=None, seed=8, chunk=1 << 20):
    """Strong scaling of ONE head over `nranks` GPUs as a spatial domain decomposition: the global strand ids sorted by the Morton
    code of their roots (10 bits per axis over the reference's grid box), cut into `nranks` contiguous pieces; returns the ids of
    piece `rank` (uint64, ascending Morton order).  Contiguous id ranges would hand every rank a uniformly thinned copy of the
    whole head (the ids are hashed): fewer strands per voxel on every rank, which is what the splat's warp aggregation and the
    collider candidate mask live on."""
    if colliders is None:
        colliders = reference_colliders()
    X, _ = _head_frames(colliders)
    S = int(num_strands_total)
    keys = np.empty(S, np.uint64)
    old = np.seterr(over="ignore")
    try:
        for lo in range(0, S, chunk):
            hi = min(S, lo + chunk)
            _, root = _head_roots(np.arange(lo, hi, dtype=np.uint64), X, seed)
            q = np.clip(((root - np.array([-3.0, -2.0, -5.0], np.float32)) * np.float32(1024.0 / 7.0)).astype(np.int64), 0, 1023).astype(np.uint64)
            k = np.zeros(hi - lo, np.uint64)
            for b in range(10):
                for a in range(3):
                    k |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
            keys[lo:hi] = k
    finally:
        np.seterr(**old)
    order = np.argsort(keys, kind="stable").astype(np.uint64)
    lo, hi = shard_range(S, rank, nranks)
    return order[lo:hi]


def synthetic_head(num_strands, num_points, strand_length=2.5, first_strand=0, colliders=None, seed=8,
                   chunk=1 << 18, out=None, ids=None):
    """Strand[S] AoS (float32 [S][3][N][4]) for global strand ids first_strand .. first_strand+S-1, or for the ids in `ids`."""
    if colliders is None:
        colliders = reference_colliders()
    X, IT = _head_frames(colliders)
    S, N = int(num_strands), int(num_points)
    rest = np.float32(np.float32(strand_length) / np.float32(N - 1))
    if out is None:
        out = np.empty((S, 3, N, 4), np.float32)
    j = np.arange(N, dtype=np.float32)
    old = np.seterr(over="ignore")
    try:
        for lo in range(0, S, chunk):
            hi = min(S, lo + chunk)
            sid = (np.arange(lo, hi, dtype=np.uint64) + np.uint64(first_strand)) if ids is None else np.asarray(ids[lo:hi], np.uint64)
            u, root = _head_roots(sid, X, seed)
            n = (u @ IT[:3, :3].T).astype(np.float32)
            n /= np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
            d = n + np.float32(0.1) * np.array([0.05, 5.0, -2.0], np.float32)
            d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
            blk = out[lo:hi]
            blk[:, 0, :, :3] = root[:, None, :] + (j * rest)[None, :, None] * d[:, None, :]
            blk[:, 0, :, 3] = 1.0
            blk[:, 1, :, :] = np.array([0.0, 0.0, -1.0, 0.0], np.float32)
            blk[:, 2, :, :] = 0.0
    finally:
        np.seterr(**old)
    return out


def triangle_cdf(tri_pos):
    """Area CDF of a triangle soup [ntris, 3, 3] as rvh_init_from_mesh builds it: double areas, sequential sum, float32."""
    t = np.asarray(tri_pos, np.float32).reshape(-1, 3, 3).astype(np.float64)
    c = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    area = 0.5 * np.sqrt(c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])
    acc = np.cumsum(area)
    cdf = (acc / acc[-1]).astype(np.float32)
    cdf[-1] = 1.0
    return cdf


def mesh_head(num_strands, num_points, strand_length, tri_pos, tri_nrm=None, first_strand=0, seed=8):
    """Host twin of rvh_init_from_mesh (k_mesh_follicles_aos): area-weighted follicles on a triangle soup."""
    tp = np.asarray(tri_pos, np.float32).reshape(-1, 3, 3)
    cdf = triangle_cdf(tp)
    S, N = int(num_strands), int(num_points)
    rest = np.float32(np.float32(strand_length) / np.float32(N - 1))
    old = np.seterr(over="ignore")
    try:
        sid = np.arange(S, dtype=np.uint64) + np.uint64(first_strand)
        base = (np.uint64(seed) << np.uint64(32)) + np.uint64(3) * sid
        r = [(_splitmix64(base + np.uint64(k)) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24) for k in range(3)]
    finally:
        np.seterr(**old)
    tri = np.searchsorted(cdf, r[0], side="right")            # smallest t with cdf[t] > r0
    tri = np.minimum(tri, len(cdf) - 1)
    u, v = r[1].copy(), r[2].copy()
    fold = (u + v) >= np.float32(1.0)
    u[fold] = np.float32(1.0) - u[fold]
    v[fold] = np.float32(1.0) - v[fold]
    w = (np.float32(1.0) - u) - v
    A, B, Cc = tp[tri, 0], tp[tri, 1], tp[tri, 2]
    root = (A * w[:, None] + B * u[:, None]) + Cc * v[:, None]
    if tri_nrm is not None:
        tn = np.asarray(tri_nrm, np.float32).reshape(-1, 3, 3)
        n = (tn[tri, 0] * w[:, None] + tn[tri, 1] * u[:, None]) + tn[tri, 2] * v[:, None]
    else:
        n = np.cross(B - A, Cc - A).astype(np.float32)
    n = n / np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    d = n + np.float32(0.1) * np.array([0.05, 5.0, -2.0], np.float32)
    d = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    out = np.zeros((S, 3, N, 4), np.float32)
    j = np.arange(N, dtype=np.float32)
    out[:, 0, :, :3] = root[:, None, :] + (j * rest)[None, :, None] * d[:, None, :]
    out[:, 0, :, 3] = 1.0
    out[:, 1, :, :] = np.array([0.0, 0.0, -1.0, 0.0], np.float32)
    return out, tri


def shard_range(num_strands, rank, nranks):
    """Contiguous strand range [lo, hi) owned by `rank`."""
    lo = (num_strands * rank) // nranks
    hi = (num_strands * (rank + 1)) // nranks
    return lo, hi
