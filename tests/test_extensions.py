"""North-star extensions that the reference does not contain (SURVEY.md top table, section 8 f3): head-SDF collision
(TMA-staged on the GPU) and density-gradient repulsion.  Each has its own oracle mode (oracle.c ORC_SDF_ON /
ORC_REPULSION_ON), which is the definition the CUDA path is held to; the analytic-ellipsoid path stays the
reference-parity path (tests/test_parity_gpu.py).

CPU tests pin the oracle modes against closed forms; GPU tests (-m gpu) compare the CUDA path, through the C ABI,
with the oracle on the same inputs -- per-step protocol and tolerances as in test_parity_gpu.py.
"""
import numpy as np
import pytest

import orc
import rvh_b200 as rvh

DT = np.float32(1.0 / 60.0)
gpu = pytest.mark.gpu

# lattice used by most tests: covers the head, neck and shoulders of the shipped scene (main.cpp:229-237)
SDF_ORIGIN = np.array([-2.0, -2.2, -1.8], np.float32)


def sdf_lattice(cell):
    ext = np.array([4.0, 6.2, 3.4], np.float32)
    dim = [int(np.ceil(e / cell)) + 1 for e in ext]
    return dim, SDF_ORIGIN, np.float32(cell)


def uv_sphere(radius, centre, nlat=24, nlon=48):
    """Closed, outward-oriented triangle mesh of a sphere."""
    verts = [[0.0, radius, 0.0]]
    for i in range(1, nlat):
        th = np.pi * i / nlat
        for j in range(nlon):
            ph = 2 * np.pi * j / nlon
            verts.append([radius * np.sin(th) * np.cos(ph), radius * np.cos(th), radius * np.sin(th) * np.sin(ph)])
    verts.append([0.0, -radius, 0.0])
    verts = np.array(verts, np.float32) + np.asarray(centre, np.float32)
    tris = []
    ring = lambda i, j: 1 + (i - 1) * nlon + (j % nlon)
    south = len(verts) - 1
    for j in range(nlon):
        tris.append([0, ring(1, j + 1), ring(1, j)])
        tris.append([south, ring(nlat - 1, j), ring(nlat - 1, j + 1)])
    for i in range(1, nlat - 1):
        for j in range(nlon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            tris.append([a, b, c])
            tris.append([b, d, c])
    return verts, np.array(tris, np.int32)


def synth(S, N, L, seed_vel=0):
    st = rvh.scenes.synthetic_head(S, N, L)
    if seed_vel:
        rng = np.random.default_rng(seed_vel)
        st[:, 1, 1:, :3] += rng.normal(scale=0.3, size=(S, N - 1, 3)).astype(np.float32)
    return st


# ---- CPU: the oracle modes against closed forms ------------------------------------------------------------

def test_oracle_collider_bake_is_exact_for_a_sphere_shaped_ellipsoid():
    cols = np.stack([orc.collider_build([9, 9, 9], [0, 0, 0], [1, 1, 1]), orc.collider_build([0.3, 0.1, -0.2], [0, 0, 0], [0.7, 0.7, 0.7])])
    dim, origin, cell = [24, 20, 22], np.array([-1.0, -0.9, -1.2], np.float32), np.float32(0.1)
    vol = orc.sdf_bake_colliders(cols, dim, origin, cell)
    k, j, i = np.meshgrid(np.arange(dim[2]), np.arange(dim[1]), np.arange(dim[0]), indexing="ij")
    p = origin + cell * np.stack([i, j, k], -1).astype(np.float32)
    exact = np.linalg.norm(p - np.array([0.3, 0.1, -0.2], np.float32), axis=-1) - 0.7
    assert np.abs(vol - exact).max() < 2e-6
    assert (vol < 0).sum() > 100


def test_oracle_mesh_bake_matches_the_analytic_sphere():
    verts, tris = uv_sphere(0.8, [0.1, 0.0, -0.1])
    dim, origin, cell = [22, 22, 22], np.array([-1.0, -1.05, -1.1], np.float32), np.float32(0.1)
    vol = orc.sdf_bake_mesh(verts, tris, dim, origin, cell)
    k, j, i = np.meshgrid(np.arange(dim[2]), np.arange(dim[1]), np.arange(dim[0]), indexing="ij")
    p = origin + cell * np.stack([i, j, k], -1).astype(np.float32)
    exact = np.linalg.norm(p - np.array([0.1, 0.0, -0.1], np.float32), axis=-1) - 0.8
    assert np.abs(vol - exact).max() < 0.012                   # faceting error of a 24 x 48 sphere
    clear = np.abs(exact) > 0.02
    assert np.array_equal(vol[clear] < 0, exact[clear] < 0)   # winding-number sign


def test_oracle_mesh_bake_sign_survives_an_open_mesh():
    """The mannequin's neck is open; the generalized winding number still separates inside from outside."""
    verts, tris = uv_sphere(0.8, [0, 0, 0])
    keep = verts[tris].mean(axis=1)[:, 1] > -0.72             # cut a hole around the south pole
    vol = orc.sdf_bake_mesh(verts, tris[keep], [5, 5, 5], np.array([-0.3, -0.2, -0.3], np.float32), np.float32(0.15))
    assert (vol < 0).all()                                    # every node is well inside
    far = orc.sdf_bake_mesh(verts, tris[keep], [3, 3, 3], np.array([1.5, 1.5, 1.5], np.float32), np.float32(0.2))
    assert (far > 0).all()


def test_oracle_sdf_sample_is_trilinear_and_bounded():
    rng = np.random.default_rng(0)
    vol = rng.normal(size=(6, 5, 7)).astype(np.float32)
    origin, cell = np.array([1.0, -2.0, 0.5], np.float32), np.float32(0.25)
    orc.set_head_sdf(vol, origin, cell)
    try:
        for (i, j, k) in [(0, 0, 0), (3, 2, 4), (5, 3, 4)]:
            ok, d, g = orc.sdf_sample(origin + cell * np.array([i, j, k], np.float32))
            assert ok and d == vol[k, j, i]                    # nodes are reproduced exactly
        ok, d, g = orc.sdf_sample(origin + cell * np.array([2.5, 1.5, 3.5], np.float32))
        assert ok and abs(d - vol[3:5, 1:3, 2:4].mean()) < 1e-6
        assert not orc.sdf_sample(origin + cell * np.array([6.0, 1.0, 1.0], np.float32))[0]    # last node: no cell
        assert not orc.sdf_sample(origin - cell)[0]
        assert not orc.sdf_sample(np.array([np.nan, 0, 0], np.float32))[0]
    finally:
        orc.set_head_sdf(None, None, 0)


def test_oracle_sdf_collision_tracks_the_analytic_ellipsoids():
    """Sanity of the semantics: with the volume baked from the ellipsoids themselves on a fine lattice, the SDF step
    lands where the analytic step lands (to lattice accuracy, not to the reference tolerance: SURVEY.md section 7)."""
    S, N, L = 3000, 16, 2.5
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, L, seed_vel=2)
    rest = np.float32(L) / np.float32(N - 1)
    pa = orc.default_params(S, N, 0, rest_length=rest)
    ps = orc.default_params(S, N, orc.SDF_ON, rest_length=rest)
    dim, origin, cell = sdf_lattice(0.04)
    vol = orc.sdf_bake_colliders(cols, dim, origin, cell)
    orc.set_head_sdf(vol, origin, cell)
    try:
        a, s = st, st
        for k in range(4):
            a, _ = orc.step(pa, cols, DT, 0.0, a)
            s, _ = orc.step(ps, cols, DT, 0.0, s)
        diff = np.abs(a[:, 0, :, :3] - s[:, 0, :, :3]).max()
        moved = np.abs(a[:, 0, :, :3] - st[:, 0, :, :3]).max()
        assert moved > 0.05 and diff < 0.1 * moved, (diff, moved)
        seg = np.linalg.norm(s[:, 0, 1:, :3] - s[:, 0, :-1, :3], axis=2)
        assert np.abs(seg / rest - 1).max() < 1e-5            # FTL invariants hold in the extension too
        assert np.array_equal(s[:, 0, 0], st[:, 0, 0])
    finally:
        orc.set_head_sdf(None, None, 0)


def test_oracle_repulsion_pushes_down_the_density_gradient_and_vanishes_at_zero():
    S, N, L = 4000, 8, 0.4
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, L, seed_vel=4)
    rest = np.float32(L) / np.float32(N - 1)
    p0 = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    p1 = orc.default_params(S, N, orc.GRID_ON | orc.REPULSION_ON, rest_length=rest)
    a, ga = orc.step(p0, cols, DT, 0.0, st)
    b, gb = orc.step(p1, cols, DT, 0.0, st)
    assert np.array_equal(ga, gb) and np.array_equal(a[:, 0], b[:, 0])      # positions and splat are untouched
    dv = b[:, 1, 1:, :3] - a[:, 1, 1:, :3]
    assert np.abs(dv).max() > 1e-3
    # the fur patch is densest in its middle: on average the push points away from the centroid
    c = a[:, 0, 1:, :3].reshape(-1, 3).mean(axis=0)
    out = a[:, 0, 1:, :3] - c
    assert (dv * out).sum() > 0
    p1.repulsion = 0.0
    z, _ = orc.step(p1, cols, DT, 0.0, st)
    assert np.array_equal(z.view(np.uint32), a.view(np.uint32))


# ---- GPU: the CUDA path against the oracle modes -----------------------------------------------------------

def _sim(S, N, flags, rest, cols, spt=0, repulsion=None):
    cfg = rvh.default_config(S, N, flags=flags, rest_length=float(rest), strands_per_thread=spt, repulsion=repulsion)
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    return sim


@gpu
def test_gpu_collider_bake_matches_oracle():
    cols = rvh.scenes.bench_colliders()
    dim, origin, cell = sdf_lattice(0.08)
    sim = _sim(256, 4, 0, 0.1, cols)
    sim.bake_head_sdf_from_colliders(dim, origin, cell)
    got = sim.download_head_sdf()
    sim.close()
    want = orc.sdf_bake_colliders(cols, dim, origin, cell)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5
    assert (got < 0).sum() > 1000


@gpu
def test_gpu_mesh_bake_matches_oracle():
    verts, tris = uv_sphere(0.8, [0.1, 0.0, -0.1])
    dim, origin, cell = [31, 29, 30], np.array([-1.1, -1.05, -1.2], np.float32), np.float32(0.075)
    sim = _sim(256, 4, 0, 0.1, rvh.scenes.bench_colliders())
    sim.bake_head_sdf_from_mesh(verts, tris, dim, origin, cell)
    got = sim.download_head_sdf()
    sim.close()
    want = orc.sdf_bake_mesh(verts, tris, dim, origin, cell)
    # distances agree to rounding; the sign can only differ where the winding number sits on 1/2, i.e. on the surface
    same = np.sign(got) == np.sign(want)
    assert np.abs(np.abs(got) - np.abs(want)).max() < 1e-5
    assert np.abs(want[~same]).max(initial=0.0) < 1e-3
    assert same.mean() > 0.999 and (got < 0).sum() > 500


@gpu
@pytest.mark.parametrize("S,N,L,extra,spt", [
    (6000, 32, 2.5, 0, 0),
    (6000, 32, 2.5, rvh.WIND_B | rvh.GRID_ON, 2),
    (70000, 16, 2.5, rvh.GRID_ON, 2),        # two strands per thread (the default from 128K strands on)
    (3001, 24, 2.5, 0, 1),
    (900, 10, 2.5, rvh.GRID_ON, 0),
])
def test_gpu_sdf_step_matches_oracle_and_tma_equals_plain_loads(S, N, L, extra, spt):
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    dim, origin, cell = sdf_lattice(0.05)
    vol = orc.sdf_bake_colliders(cols, dim, origin, cell)
    oflags = orc.SDF_ON | (orc.GRID_ON if extra & rvh.GRID_ON else 0) | (orc.WIND_B if extra & rvh.WIND_B else 0)
    p = orc.default_params(S, N, oflags, rest_length=rest)
    orc.set_head_sdf(vol, origin, cell)
    try:
        sims = {}
        for name, f in (("tma", rvh.SDF_TMA), ("ldg", 0)):
            sims[name] = _sim(S, N, rvh.SDF_ON | extra | f, rest, cols, spt=spt)
            sims[name].set_head_sdf(vol, origin, cell)
        assert sims["tma"].sdf_mode() == "tma" and sims["ldg"].sdf_mode() == "ldg"
        assert np.array_equal(sims["tma"].download_head_sdf(), vol)
        state = synth(S, N, L, seed_vel=3)
        for k in range(20):                                     # let the hair fall onto the head first (oracle free run)
            state, _ = orc.step(p, cols, DT, 0.3, state, threads=8)
        hits = 0
        rng = np.random.default_rng(5)
        for k in range(4):                                      # resynchronise every step
            probe = state[:, 0, 1:, :3].reshape(-1, 3)[rng.integers(0, S * (N - 1), 1500)]
            hits += sum(1 for ok, d, _ in (orc.sdf_sample(q) for q in probe) if ok and d < 0)
            ref, _ = orc.step(p, cols, DT, 0.3, state, threads=8)
            outs = {}
            for name, sim in sims.items():
                sim.upload(state)
                sim.step(DT, 0.3)
                outs[name] = sim.download()
            assert np.array_equal(outs["tma"].view(np.uint32), outs["ldg"].view(np.uint32)), "TMA-staged tiles and plain loads must agree bit for bit"
            perr = np.abs(outs["tma"][:, 0, :, :3] - ref[:, 0, :, :3]).max()
            verr = np.abs(outs["tma"][:, 1, :, :3] - ref[:, 1, :, :3]).max()
            assert perr <= 1e-4 * L, "step %d position error %.3e" % (k, perr)
            assert verr <= 1e-4 * L / float(DT), "step %d velocity error %.3e" % (k, verr)
            seg = np.linalg.norm(outs["tma"][:, 0, 1:, :3].astype(np.float64) - outs["tma"][:, 0, :-1, :3], axis=2)
            assert np.abs(seg / rest - 1).max() < 2e-5
            assert np.array_equal(outs["tma"][:, 0, 0], state[:, 0, 0])
            state = ref
        assert hits > 20, "the scene must actually exercise the SDF collision (%d of 6000 probes inside)" % hits
        for sim in sims.values():
            sim.close()
    finally:
        orc.set_head_sdf(None, None, 0)


@gpu
def test_gpu_sdf_free_running_steps_use_the_tile_pipeline_every_row():
    """Several steps back to back (fused gather + SDF tiles), against free-running plain loads: bit-identical."""
    S, N, L = 20000, 48, 2.5
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    dim, origin, cell = sdf_lattice(0.05)
    vol = orc.sdf_bake_colliders(cols, dim, origin, cell)
    st = synth(S, N, L, seed_vel=9)
    outs = []
    for f in (rvh.SDF_TMA, 0):
        sim = _sim(S, N, rvh.SDF_ON | rvh.GRID_ON | rvh.WIND_B | f, rest, cols)
        sim.set_head_sdf(vol, origin, cell)
        sim.upload(st)
        for k in range(6):
            sim.step(DT, 0.1 * k)
        outs.append(sim.download())
        sim.close()
    assert np.isfinite(outs[0]).all()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


@gpu
def test_gpu_sdf_requires_a_volume():
    sim = _sim(64, 8, rvh.SDF_ON, 0.3, rvh.scenes.bench_colliders())
    sim.upload(synth(64, 8, 2.5))
    with pytest.raises(rvh.RvhError):
        sim.step(DT)
    with pytest.raises(rvh.RvhError):
        sim.set_head_sdf(np.zeros((1, 4, 4), np.float32), [0, 0, 0], 0.1)     # needs >= 2 nodes per axis
    sim.close()
    with pytest.raises(rvh.RvhError):
        rvh.HairSim(rvh.default_config(64, 8, flags=rvh.REPULSION_ON))          # repulsion reads the grid


@gpu
@pytest.mark.parametrize("S,N,L,spt", [(20000, 16, 0.4, 0), (5000, 32, 2.5, 2), (70000, 8, 0.4, 0)])
def test_gpu_repulsion_matches_oracle(S, N, L, spt):
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    st = synth(S, N, L, seed_vel=6)
    p = orc.default_params(S, N, orc.GRID_ON | orc.REPULSION_ON, rest_length=rest)
    p0 = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    sim = _sim(S, N, rvh.GRID_ON | rvh.REPULSION_ON, rest, cols, spt=spt)
    state = st
    for k in range(3):
        ref, ref_grid = orc.step(p, cols, DT, 0.0, state, threads=8)
        plain, _ = orc.step(p0, cols, DT, 0.0, state, threads=8)
        sim.upload(state)
        sim.step(DT, 0.0)
        out = sim.download()                                   # stand-alone k_grid_gather<REP>
        assert abs(int(sim.download_grid()[:, 3].sum()) - int(ref_grid[:, 3].sum())) <= 8 * S * N      # same density up to rounding flips
        assert np.abs(out[:, 0, :, :3] - ref[:, 0, :, :3]).max() <= 1e-4 * L
        dv_ref = ref[:, 1, :, :3] - plain[:, 1, :, :3]         # what the repulsion added
        dv_gpu = out[:, 1, :, :3] - plain[:, 1, :, :3]
        assert 1e-3 < np.abs(dv_ref).max() <= 0.2 * 1.0001       # every component is bounded by `repulsion`
        # the gradient of a trilinear field jumps across cell faces: points within 1e-4 cell of a face may sit on either
        # side of it in the two implementations (positions agree to ~1e-7) and are excluded from the tight comparison
        g = (ref[:, 0, :, :3] - np.array([-3, -2, -5], np.float32)) / (np.float32(7.0) / np.float32(64.0))
        clear = (np.abs(g - np.round(g)) > 1e-4).all(axis=2)
        assert clear.mean() > 0.99
        assert np.abs((dv_gpu - dv_ref)[clear]).max() <= 1e-4
        assert np.abs(dv_gpu - dv_ref).max() <= 2 * 0.2          # ... and even on a face the jump is bounded
        state = ref
    sim.close()


@gpu
@pytest.mark.parametrize("flags", [rvh.GRID_ON, rvh.GRID_ON | rvh.REPULSION_ON])
def test_gpu_fused_gather_equals_stand_alone_gather(flags):
    """The gather of step k normally rides in k_ftl_step of step k+1; reading the state back in between runs the
    stand-alone kernel instead.  Same arithmetic either way."""
    S, N, L = 30000, 16, 0.4
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    st = synth(S, N, L, seed_vel=8)
    a = _sim(S, N, flags, rest, cols)
    a.upload(st)
    a.step(DT, 0.0)
    a.step(DT, 0.1)                                            # fused gather of step 0's grid
    fused = a.download()
    a.close()
    b = _sim(S, N, flags, rest, cols)
    b.upload(st)
    b.step(DT, 0.0)
    mid = b.download()                                         # forces k_grid_gather
    b.upload(mid)
    b.step(DT, 0.1)
    split = b.download()
    b.close()
    assert np.abs(fused[:, 0, :, :3] - split[:, 0, :, :3]).max() <= 1e-6
    assert np.abs(fused[:, 1, :, :3] - split[:, 1, :, :3]).max() <= 1e-4


# ---- guide strand -> render strands (hair.tesc / hair.tese) -------------------------

def test_oracle_expansion_follows_hair_tese():
    """Closed-form properties of the restated tessellation-evaluation shader."""
    S, N = 40, 10
    st = synth(S, N, 2.5, seed_vel=1)
    rng = np.random.default_rng(2)
    st[:, 0, 1:, :3] += rng.normal(scale=0.02, size=(S, N - 1, 3)).astype(np.float32)      # not straight
    pw, tu = orc.expand_strands(st, 12, 42)
    assert pw.shape == (S, 12, 43, 4)
    # isoline coordinate and strand width (hair.tese:313-315)
    assert np.allclose(tu[..., 3], (np.arange(12) / 12.0)[None, :, None])
    assert np.allclose(pw[..., 3], 0.02 + (np.arange(43) / 42.0) * (0.01 - 0.02), atol=1e-7)
    # the sideways deviation is horizontal (dir.y = 0) and vanishes nowhere but is bounded by width * sd_max
    # => y follows the Bezier curve exactly, and at the segment joints the curve passes through the guide points
    joints = [(j, j * 9 // 42) for j in range(43) if (j * 9) % 42 == 0]
    for j, i in joints:
        assert np.abs(pw[:, :, j, 1] - st[:, 0, i, 1][:, None]).max() < 2e-6
    # v = 0: sd = 1 (hair.tese:267-269) => offset = width(0) * (rand2 + 0.5) along dir(u), same for every strand
    off = pw[:, :, 0, :3] - st[:, 0, 0, :3][:, None, :]
    assert np.abs(off - off[0:1]).max() < 1e-6
    r = np.linalg.norm(off[0], axis=1)
    assert np.all(r >= 0.5 * 0.05 * 0.5 - 1e-6) and np.all(r <= 0.5 * 0.05 * 1.5 + 1e-6)
    ang = np.arctan2(off[0, :, 2], off[0, :, 0]) % (2 * np.pi)
    assert np.allclose(ang, 2 * np.pi * np.arange(12) / 12.0, atol=1e-4)
    # unit tangents of the guide's own segments
    assert np.abs(np.linalg.norm(tu[..., :3], axis=-1) - 1).max() < 1e-6
    seg = np.minimum((np.arange(43) / np.float32(42.0) * 9).astype(int), 8)
    d = st[:, 0, seg + 1, :3] - st[:, 0, seg, :3]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    assert np.abs(tu[:, 0, :, :3] - d).max() < 1e-6
    # all five deviation profiles and "none" occur across strands x isolines
    dev = np.linalg.norm(pw[:, :, -1, [0, 2]] - pw[:, :, 21, [0, 2]], axis=-1)
    assert len(np.unique(np.round(dev / (dev.max() + 1e-9), 2))) > 6


@gpu
@pytest.mark.parametrize("S,N,I,D", [(900, 10, 12, 42), (5000, 32, 12, 42), (1500, 16, 5, 17), (33, 2, 3, 4)])
def test_gpu_expansion_matches_oracle(S, N, I, D):
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(2.5) / np.float32(N - 1)
    st = synth(S, N, 2.5, seed_vel=5)
    sim = _sim(S, N, rvh.GRID_ON, rest, cols)
    sim.upload(st)
    for k in range(3):
        sim.step(DT, 0.0)
    state = sim.download()
    pw, tu, ms = sim.expand(I, D)
    sim.close()
    ref_pw, ref_tu = orc.expand_strands(state, I, D)
    assert pw.shape == ref_pw.shape == (S, I, D + 1, 4)
    assert np.abs(pw - ref_pw).max() <= 2e-6 * 4.0          # FMA contraction only: every transcendental is tabulated on the host
    assert np.abs(tu - ref_tu).max() <= 2e-6
    assert ms > 0


# ---- GPU follicle placement on a mesh, area weighted (Strand.cpp:92 TODO) ----------------

def scalp_soup():
    """The reference's follicle surface (models/mannequin_segment.obj, frozen as arrays under tests/golden/): quads -> triangle soup."""
    import os
    m = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mannequin_segment_mesh.npz"))
    fv, fn = m["fv"], m["fn"]
    idx = np.concatenate([fv[:, [0, 1, 2]], fv[:, [0, 2, 3]]])
    nidx = np.concatenate([fn[:, [0, 1, 2]], fn[:, [0, 2, 3]]])
    return m["v"][idx].astype(np.float32), m["vn"][nidx].astype(np.float32)


def test_host_mesh_follicles_are_area_weighted_and_on_the_surface():
    tp, tn = scalp_soup()
    S = 200000
    st, tri = rvh.scenes.mesh_head(S, 4, 0.3, tp, tn)
    area = 0.5 * np.linalg.norm(np.cross(tp[:, 1] - tp[:, 0], tp[:, 2] - tp[:, 0]), axis=1)
    want = area / area.sum() * S
    got = np.bincount(tri, minlength=len(tp))
    big = want > 50
    assert np.abs(got[big] - want[big]).max() < 6 * np.sqrt(want[big]).max()            # Poisson scatter, not uniform-per-triangle
    assert np.corrcoef(got, want)[0, 1] > 0.95
    # roots lie in their triangle's plane
    nrm = np.cross(tp[:, 1] - tp[:, 0], tp[:, 2] - tp[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    off = np.einsum("ij,ij->i", st[:, 0, 0, :3] - tp[tri, 0], nrm[tri])
    assert np.abs(off).max() < 1e-5
    # shards reproduce the global sequence
    part, _ = rvh.scenes.mesh_head(1000, 4, 0.3, tp, tn, first_strand=5000)
    assert np.array_equal(part, st[5000:6000])


@gpu
def test_gpu_mesh_follicles_match_host_twin():
    tp, tn = scalp_soup()
    S, N, L = 20000, 12, 1.0
    for first, normals in ((0, tn), (777, None)):
        sim = _sim(S, N, 0, np.float32(L) / np.float32(N - 1), rvh.scenes.bench_colliders())
        sim.init_from_mesh(tp, normals, first_strand=first, strand_length=L, seed=8)
        got = sim.download()
        sim.close()
        want, _ = rvh.scenes.mesh_head(S, N, L, tp, normals, first_strand=first, seed=8)
        assert np.abs(got[:, 0, :, :3] - want[:, 0, :, :3]).max() <= 4e-6 * 4.0
        assert np.array_equal(got[:, 1], want[:, 1]) and np.all(got[:, 2] == 0) and np.all(got[:, 0, :, 3] == 1)


# ---- the real head: SDF baked from the mannequin mesh (reference asset models/mannequin.obj, main.cpp:222-223) ----

def head_mesh():
    import os
    m = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mannequin_head_mesh.npz"))
    return (m["v"] * np.float32(0.98)).astype(np.float32), m["tri"]          # rendered with glm::scale(0.98), main.cpp:223


HEAD_LATTICE = ([30, 55, 24], np.array([-1.45, -1.6, -1.15], np.float32), np.float32(0.1))


def test_oracle_head_mesh_bake_is_a_plausible_signed_distance():
    verts, tris = head_mesh()
    dim, origin, cell = HEAD_LATTICE
    vol = orc.sdf_bake_mesh(verts, tris, dim, origin, cell)
    assert np.isfinite(vol).all()
    inside = vol < 0
    assert 0.08 < inside.mean() < 0.5                              # the head, neck and bust fill part of the box
    # the centre of the head ellipsoid of the shipped scene (main.cpp:231) is inside, a corner of the box is outside
    c = np.round((np.array([0.0, 2.64, 0.08], np.float32) - origin) / cell).astype(int)
    assert vol[c[2], c[1], c[0]] < -0.3
    assert vol[0, 0, 0] > 0.3 and vol[-1, -1, -1] > 0.1
    # |grad d| ~ 1 away from the medial axis: central differences over the lattice
    g = np.stack(np.gradient(vol, float(cell)), -1)
    gn = np.linalg.norm(g, axis=-1)
    assert 0.8 < np.median(gn) < 1.1


@gpu
def test_gpu_head_mesh_bake_and_step_match_oracle():
    verts, tris = head_mesh()
    dim, origin, cell = HEAD_LATTICE
    want = orc.sdf_bake_mesh(verts, tris, dim, origin, cell)
    S, N, L = 8000, 24, 2.5
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    sim = _sim(S, N, rvh.SDF_ON | rvh.GRID_ON, rest, cols)
    sim.bake_head_sdf_from_mesh(verts, tris, dim, origin, cell)
    got = sim.download_head_sdf()
    same = np.sign(got) == np.sign(want)
    assert np.abs(np.abs(got) - np.abs(want)).max() < 2e-5
    assert same.mean() > 0.999 and np.abs(want[~same]).max(initial=0.0) < 2e-3
    # one step against the oracle, both sampling the GPU-baked volume
    p = orc.default_params(S, N, orc.SDF_ON | orc.GRID_ON, rest_length=rest)
    orc.set_head_sdf(got, origin, cell)
    try:
        state = synth(S, N, L, seed_vel=4)
        for k in range(15):
            state, _ = orc.step(p, cols, DT, 0.0, state, threads=8)
        probe = state[::5, 0, 1:, :3].reshape(-1, 3)[:3000]
        assert sum(1 for ok, d, _ in (orc.sdf_sample(q) for q in probe) if ok and d < 0) > 10
        ref, _ = orc.step(p, cols, DT, 0.0, state, threads=8)
        sim.upload(state)
        sim.step(DT, 0.0)
        out = sim.download()
        assert np.abs(out[:, 0, :, :3] - ref[:, 0, :, :3]).max() <= 1e-4 * L
        assert np.abs(out[:, 1, :, :3] - ref[:, 1, :, :3]).max() <= 1e-4 * L / float(DT)
    finally:
        orc.set_head_sdf(None, None, 0)
        sim.close()


# ---- frozen vectors of the extension modes (tests/golden/make_extension_golden.py) -------------------------------

def _golden_ext():
    import os
    from conftest import full_state
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extensions.npz")))
    g["sdf"] = orc.sdf_bake_colliders(g["colliders"], g["sdf_dim"], g["sdf_origin"], g["sdf_cell"])      # re-baked; pinned by slice + sum
    return g, full_state(g["pre"])


GOLDEN_MODES = [("post_sdf", orc.SDF_ON | orc.GRID_ON), ("post_rep", orc.GRID_ON | orc.REPULSION_ON), ("post_both", orc.SDF_ON | orc.GRID_ON | orc.REPULSION_ON)]


def test_oracle_extension_modes_reproduce_the_frozen_vectors():
    g, pre = _golden_ext()
    S, _, N, _ = pre.shape
    rest = np.float32(2.5) / np.float32(N - 1)
    assert np.abs(g["sdf"][int(g["sdf_dim"][2]) // 2] - g["sdf_slice"]).max() <= 1e-6
    assert abs(float(g["sdf"].astype(np.float64).sum()) - float(g["sdf_sum"])) <= 1e-3
    orc.set_head_sdf(g["sdf"], g["sdf_origin"], g["sdf_cell"])
    try:
        for name, fl in GOLDEN_MODES:
            out, _ = orc.step(orc.default_params(S, N, fl, rest_length=rest), g["colliders"], DT, 0.0, pre)
            assert np.abs(out[:, 0:2, :, :3] - g[name]).max() <= 1e-6, name       # same libm => normally bit-identical
        assert np.abs(g["post_sdf"] - g["post_rep"]).max() > 1e-3                # the modes really differ
    finally:
        orc.set_head_sdf(None, None, 0)
    pw, tu = orc.expand_strands(pre, 4, 9)
    assert np.abs(pw - g["exp_pw"]).max() <= 1e-6 and np.abs(tu - g["exp_tu"]).max() <= 1e-6


@gpu
def test_gpu_extension_modes_against_the_frozen_vectors():
    g, pre = _golden_ext()
    S, _, N, _ = pre.shape
    L = 2.5
    rest = np.float32(L) / np.float32(N - 1)
    for name, fl in GOLDEN_MODES:
        flags = (rvh.SDF_ON if fl & orc.SDF_ON else 0) | (rvh.GRID_ON if fl & orc.GRID_ON else 0) | (rvh.REPULSION_ON if fl & orc.REPULSION_ON else 0)
        sim = _sim(S, N, flags, rest, g["colliders"])
        if flags & rvh.SDF_ON:
            sim.set_head_sdf(g["sdf"], g["sdf_origin"], float(g["sdf_cell"]))
        sim.upload(pre)
        sim.step(DT, 0.0)
        out = sim.download()
        if name == "post_sdf":
            pw, tu, _ = sim.expand(4, 9)          # expansion reads positions only; compare on a fresh upload below
        sim.close()
        assert np.abs(out[:, 0, :, :3] - g[name][:, 0]).max() <= 1e-4 * L, name
        if not (fl & orc.REPULSION_ON):            # repulsion is discontinuous across cell faces (see the repulsion test)
            assert np.abs(out[:, 1, :, :3] - g[name][:, 1]).max() <= 1e-4 * L / float(DT), name
    sim = _sim(S, N, 0, rest, g["colliders"])
    sim.upload(pre)
    pw, tu, _ = sim.expand(4, 9)
    sim.close()
    assert np.abs(pw - g["exp_pw"]).max() <= 8e-6 and np.abs(tu - g["exp_tu"]).max() <= 2e-6


@gpu
def test_gpu_expansion_matches_the_reference_tese_text():
    """k_expand_strands against the frozen outputs of hair.tese itself (tests/golden/expand_tese_n10.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "expand_tese_n10.npz"))
    S, I, D = g["state"].shape[0], int(g["isolines"]), int(g["divisions"])
    st = np.zeros((S, 3, 10, 4), np.float32)
    st[:, 0:2] = g["state"]
    sim = rvh.HairSim(rvh.default_config(S, 10, flags=0))
    sim.upload(st)
    pw, tu, _ = sim.expand(I, D)
    sim.close()
    ref = g["ref"]
    assert np.abs(pw[:, :, :D, :3] - ref[..., :3]).max() <= 2e-6 * 4.0          # FMA contraction on the GPU: a few ulp of a coordinate of magnitude <= 4
    assert np.abs(pw[:, :, :D, 3] - ref[..., 3]).max() <= 1e-7
    assert np.abs(tu[:, :, :D, :3] - ref[..., 4:7]).max() <= 2e-6
    assert np.array_equal(tu[:, :, :D, 3], ref[..., 7])
