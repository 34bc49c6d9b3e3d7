"""GPU parity tests: the CUDA path (through the C ABI, librvh.so) against the CPU oracle and
the golden vectors produced by the reference's own sources.

Tolerances (BASELINE.json north_star): per step, max |position error| <= 1e-4 * strand length;
segment length within 1e-5 relative of rest length; indexing bit-exact; integer grid exact when
fed identical inputs.  "Per step" is literal: upload state k, run ONE step, compare with k+1.
"""
import numpy as np
import pytest

import orc
import rvh_b200 as rvh
from conftest import full_state

pytestmark = pytest.mark.gpu

DT = np.float32(1.0 / 60.0)
POS_TOL_REL = 1e-4       # x strand length
SEG_TOL_REL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def gpu_step(state, colliders, flags, dt=DT, total_time=0.0, rest_length=None, phases=3, spt=0, want_grid=False):
    S, _, N, _ = state.shape
    cfg = rvh.default_config(S, N, flags=flags, rest_length=rest_length, strands_per_thread=spt)
    sim = rvh.HairSim(cfg)
    try:
        sim.set_colliders(colliders)
        sim.upload(state)
        if phases == 3:
            sim.step(dt, total_time)
        else:
            sim.step_phases(dt, total_time, phases)
        out = sim.download()
        grid = sim.download_grid() if want_grid else None
        ind = sim.draw_indirect()
    finally:
        sim.close()
    assert ind == [S, 1, 0, 0]
    return out, grid


def check_state(gpu, ref, rest, N, vel_dt=DT, what=""):
    L = rest * (N - 1)
    perr = np.abs(gpu[:, 0, :, :3] - ref[:, 0, :, :3]).max()
    verr = np.abs(gpu[:, 1, :, :3] - ref[:, 1, :, :3]).max()
    assert perr <= POS_TOL_REL * L, "%s position error %.3e > %.3e" % (what, perr, POS_TOL_REL * L)
    # velocity = position difference / dt, so its tolerance is the position tolerance / dt
    assert verr <= POS_TOL_REL * L / float(vel_dt), "%s velocity error %.3e" % (what, verr)
    seg = np.linalg.norm(gpu[:, 0, 1:, :3].astype(np.float64) - gpu[:, 0, :-1, :3], axis=2)
    seg_ref = np.linalg.norm(ref[:, 0, 1:, :3].astype(np.float64) - ref[:, 0, :-1, :3], axis=2)
    # float32 coordinates of magnitude |p| cannot resolve a length to better than a few ulp(|p|):
    # the 1e-5 relative bar is applied above that representability floor (it only binds for
    # N=64, where rest = 0.04 and 1e-5*rest = 4e-7 is about one ulp of a coordinate near 4).
    floor = 4.0 * float(np.spacing(np.float32(np.abs(ref[:, 0, :, :3]).max())))
    seg_tol = max(SEG_TOL_REL * rest, floor)
    assert np.abs(seg - rest).max() <= seg_tol, "%s segment drift %.3e" % (what, np.abs(seg / rest - 1).max())
    assert np.abs(seg - seg_ref).max() <= seg_tol, "%s segment mismatch vs oracle %.3e" % (what, np.abs(seg - seg_ref).max())
    assert np.array_equal(bits(gpu[:, 0, 0]), bits(ref[:, 0, 0])), "roots must be bit-unchanged"
    assert np.all(gpu[:, 0, :, 3] == 1.0) and np.all(gpu[:, 1, :, 3] == 0.0)
    return perr, verr


# ---- C1: the shipped scene, against golden pairs from the reference shader text -----------

@pytest.mark.parametrize("idx", range(5))
def test_c1_reference_scene_step_pairs(golden_c1, idx):
    k = int(golden_c1["k"][idx])
    pre = full_state(golden_c1["pre"][idx])
    post = full_state(golden_c1["post"][idx], golden_c1["corr_post"][idx])
    flags = rvh.GRID_ON | rvh.GRID_INT32_WRAP | rvh.KEEP_CORRECTION
    out, grid = gpu_step(pre, golden_c1["colliders"], flags, total_time=float(k) * float(DT), want_grid=True)
    rest = np.float32(2.5) / np.float32(9.0)
    # positions never depend on the grid: pinned by the reference run itself
    perr, verr = check_state(out, post, float(rest), 10, what="C1 k=%d" % k)
    cerr = np.abs(out[:, 2, :, :3] - post[:, 2, :, :3]).max()
    assert cerr <= POS_TOL_REL * 2.5
    assert np.all(out[:, 2, :, 3] == 0)
    # grid: same occupied cells up to boundary flips, totals agree closely
    ref_dens = np.zeros(64 ** 3, np.int64)
    ref_dens[golden_c1["grid_idx_%d" % k]] = golden_c1["grid_val_%d" % k][:, 3]
    assert abs(int(grid[:, 3].astype(np.int64).sum()) - int(ref_dens.sum())) <= 8 * 9 * 900
    print("C1 k=%d pos err %.2e vel err %.2e corr err %.2e" % (k, perr, verr, cerr))


# ---- synthetic heads against the live oracle ------------------------------------------------

def synth(S, N, L, seed_vel=0):
    st = rvh.scenes.synthetic_head(S, N, L)
    if seed_vel:
        rng = np.random.default_rng(seed_vel)
        st[:, 1, 1:, :3] += rng.normal(scale=0.3, size=(S, N - 1, 3)).astype(np.float32)
    return st


@pytest.mark.parametrize("S,N,L,flags", [
    (16384, 32, 2.5, rvh.WIND_B),                          # C2 shape: gravity + wind + colliders, no grid
    (4096, 32, 2.5, rvh.WIND_A),
    (5000, 64, 2.5, rvh.GRID_ON),                          # C3 shape (reduced S): grid friction, int64
    (20000, 16, 0.4, rvh.GRID_ON),                         # C4 shape (reduced S): fur
    (3000, 32, 2.5, rvh.GRID_ON | rvh.WIND_B),
    (900, 10, 2.5, 0),
])
def test_synthetic_head_one_step_vs_oracle(S, N, L, flags):
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, L, seed_vel=3)
    rest = np.float32(L) / np.float32(N - 1)
    oflags = (orc.GRID_ON if flags & rvh.GRID_ON else 0) | (orc.WIND_A if flags & rvh.WIND_A else 0) | (orc.WIND_B if flags & rvh.WIND_B else 0)
    p = orc.default_params(S, N, oflags, rest_length=rest)
    T = 0.75
    # a few steps so that points are inside colliders and the grid is busy; resynchronise each step
    state = st
    for k in range(3):
        ref, ref_grid = orc.step(p, cols, DT, T, state, threads=8)
        out, grid = gpu_step(state, cols, flags | rvh.KEEP_CORRECTION, total_time=T, rest_length=float(rest), want_grid=bool(flags & rvh.GRID_ON))
        perr, verr = check_state(out, ref, float(rest), N, what="S=%d N=%d step %d" % (S, N, k))
        if flags & rvh.GRID_ON:
            # density totals agree to within one unit per contribution
            assert abs(int(grid[:, 3].sum()) - int(ref_grid[:, 3].sum())) <= 8 * S * N
        state = ref
    inside = 0
    print("S=%d N=%d flags=%d pos err %.2e vel err %.2e" % (S, N, flags, perr, verr))


# ---- integer grid: exact when fed identical inputs -------------------------------------------

@pytest.mark.parametrize("S,N,L", [(900, 10, 2.5), (6000, 32, 2.5), (30000, 16, 0.4)])
def test_grid_splat_bit_exact_and_gather_within_ulps_given_same_inputs(S, N, L):
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, L, seed_vel=5)
    rest = np.float32(L) / np.float32(N - 1)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON, rest_length=float(rest))
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    sim.step_phases(DT, 0.0, 1)            # integrate + corrected velocity + splat, no gather
    mid = sim.download()                   # correctionVecs are zero without KEEP_CORRECTION
    grid = sim.download_grid()
    sim.step_phases(DT, 0.0, 2)            # gather only
    post = sim.download()
    sim.close()
    p = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    assert np.all(mid[:, 2] == 0)
    _, ref_grid = orc.phase_splat(p, DT, mid)     # zero corr => velocities unchanged, pure splat
    assert np.array_equal(grid, ref_grid), "grid integers differ: %d cells" % int(np.any(grid != ref_grid, axis=1).sum())
    # the gather is floating point: the CUDA path pre-divides cell velocity by density and uses FMAs, so it agrees
    # with the shader's w * (1/d) * v evaluation order to a few ulp of the cell velocity, not bit for bit
    ref_post = orc.phase_gather(p, mid, grid)
    gerr = np.abs(post[:, 1, :, :3] - ref_post[:, 1, :, :3])
    assert np.all(gerr <= 2e-6 * (np.abs(ref_post[:, 1, :, :3]) + np.abs(mid[:, 1, :, :3]) + 1.0)), "gather error %.3e" % gerr.max()
    assert np.all(post[:, 1, :, 3] == 0)
    assert np.array_equal(bits(post[:, 0]), bits(mid[:, 0]))
    assert grid[:, 3].sum() > 0


# ---- indexing / ordering / layout --------------------------------------------------------------

def test_morton_reordering_is_invisible_and_results_identical():
    S, N = 5000, 16
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, 0.4, seed_vel=7)
    rng = np.random.default_rng(1)
    st = st[rng.permutation(S)]            # external order unrelated to space
    rest = float(np.float32(0.4) / np.float32(N - 1))
    a, ga = gpu_step(st, cols, rvh.GRID_ON, rest_length=rest, want_grid=True)
    b, gb = gpu_step(st, cols, rvh.GRID_ON | rvh.KEEP_ORDER, rest_length=rest, want_grid=True)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(ga, gb)


def test_upload_download_round_trip_bit_exact():
    for S, N in [(1, 2), (31, 10), (900, 10), (1025, 7), (4097, 64)]:
        rng = np.random.default_rng(S)
        st = rng.normal(size=(S, 3, N, 4)).astype(np.float32)
        st[:, 0, :, 3] = 1
        st[:, 1, :, 3] = 0
        st[:, 2] = 0
        cfg = rvh.default_config(S, N, flags=0)
        sim = rvh.HairSim(cfg)
        sim.upload(st)
        out = sim.download()
        sim.close()
        assert np.array_equal(bits(out), bits(st)), (S, N)


@pytest.mark.parametrize("flags", [rvh.GRID_ON | rvh.WIND_B, rvh.GRID_ON | rvh.KEEP_CORRECTION, 0])
def test_step_host_equals_upload_step_download(flags):
    """rvh_step_host moves only curvePoints + curveVels over PCIe; the result must be the separate calls' result."""
    S, N = 2500, 12
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, 2.5, seed_vel=17)
    rest = float(np.float32(2.5) / np.float32(N - 1))
    ref, _ = gpu_step(st, cols, flags, total_time=0.4, rest_length=rest)
    cfg = rvh.default_config(S, N, flags=flags, rest_length=rest)
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    buf = st.copy()
    buf[:, 2] = 123.0                                   # garbage in the dead correctionVecs third of the host buffer
    sim.step_host(buf, DT, 0.4)
    sim.close()
    assert np.array_equal(bits(buf[:, 0:2]), bits(ref[:, 0:2]))
    if flags & rvh.KEEP_CORRECTION:
        assert np.array_equal(bits(buf[:, 2]), bits(ref[:, 2]))
    else:
        assert np.all(buf[:, 2] == 123.0)               # left untouched, documented in include/rvh.h


def test_gpu_scene_init_matches_host_generator():
    """rvh_init_synthetic_head (device-side splitmix64 generator) against scenes.synthetic_head (numpy), incl. a shard offset."""
    S, N, L = 5000, 16, 0.4
    cols = rvh.scenes.bench_colliders()
    for first in (0, 123457):
        cfg = rvh.default_config(S, N, flags=0, rest_length=float(np.float32(L) / np.float32(N - 1)))
        sim = rvh.HairSim(cfg)
        sim.set_colliders(cols)
        sim.init_synthetic_head(first, L, 8)
        got = sim.download()
        sim.close()
        want = rvh.scenes.synthetic_head(S, N, L, first_strand=first, colliders=cols)
        assert np.abs(got[:, 0, :, :3] - want[:, 0, :, :3]).max() <= 2e-6 * 4.0      # cos/sin/sqrt differ by an ulp or two
        assert np.array_equal(bits(got[:, 1]), bits(want[:, 1])) and np.all(got[:, 2] == 0) and np.all(got[:, 0, :, 3] == 1)


@pytest.mark.parametrize("spt", [1, 2, 4])
def test_strands_per_thread_variants_agree(spt):
    S, N = 3001, 24
    cols = rvh.scenes.bench_colliders()
    st = synth(S, N, 2.5, seed_vel=11)
    rest = np.float32(2.5) / np.float32(N - 1)
    p = orc.default_params(S, N, orc.GRID_ON | orc.WIND_B, rest_length=rest)
    ref, _ = orc.step(p, cols, DT, 0.3, st)
    out, _ = gpu_step(st, cols, rvh.GRID_ON | rvh.WIND_B, total_time=0.3, rest_length=float(rest), spt=spt)
    check_state(out, ref, float(rest), N, what="spt=%d" % spt)


# ---- edge cases ----------------------------------------------------------------------------------

@pytest.mark.parametrize("S,N", [(1, 2), (1, 10), (33, 3), (127, 10), (129, 5), (900, 10)])
def test_ragged_sizes(S, N):
    cols = rvh.scenes.reference_colliders()
    st = synth(S, N, 2.5, seed_vel=13)
    rest = np.float32(2.5) / np.float32(N - 1)
    p = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    ref, ref_grid = orc.step(p, cols, DT, 0.0, st)
    out, grid = gpu_step(st, cols, rvh.GRID_ON, rest_length=float(rest), want_grid=True)
    check_state(out, ref, float(rest), N, what="S=%d N=%d" % (S, N))


def test_points_outside_the_grid_touch_no_cell():
    S, N = 64, 8
    st = synth(S, N, 2.5)
    st[:, 0, :, :3] += np.array([100.0, -50.0, 30.0], np.float32)     # far outside [-3,4]x[-2,5]x[-5,2]
    cols = rvh.scenes.reference_colliders()
    rest = np.float32(2.5) / np.float32(N - 1)
    out, grid = gpu_step(st, cols, rvh.GRID_ON, rest_length=float(rest), want_grid=True)
    assert not grid.any()
    p = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    ref, _ = orc.step(p, cols, DT, 0.0, st)
    check_state(out, ref, float(rest), N)


def test_grid_border_cells_match_oracle():
    """Points within one cell of every face of the grid box exercise the max/min clamps (compute.comp:224-229)."""
    S, N = 600, 4
    rng = np.random.default_rng(3)
    st = np.zeros((S, 3, N, 4), np.float32)
    lo = np.array([-3, -2, -5], np.float32)
    roots = lo + rng.uniform(-0.2, 7.2, (S, 3)).astype(np.float32)
    face = rng.integers(0, 6, S)
    for s in range(S):
        ax = face[s] % 3
        roots[s, ax] = lo[ax] + (rng.uniform(-0.15, 0.15) if face[s] < 3 else 7.0 + rng.uniform(-0.15, 0.15))
    rest = np.float32(0.01)
    st[:, 0, :, :3] = roots[:, None] + (np.arange(N, dtype=np.float32) * rest)[None, :, None] * np.array([0, -1, 0], np.float32)
    st[:, 0, :, 3] = 1
    st[:, 1, :, :3] = rng.normal(size=(S, N, 3)).astype(np.float32)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON, rest_length=float(rest))
    sim = rvh.HairSim(cfg)
    sim.set_colliders(rvh.scenes.reference_colliders())
    sim.upload(st)
    sim.step_phases(DT, 0.0, 1)
    mid, grid = sim.download(), sim.download_grid()
    sim.close()
    p = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    _, ref_grid = orc.phase_splat(p, DT, mid)
    assert np.array_equal(grid, ref_grid)


def test_error_behaviour():
    cfg = rvh.default_config(64, 8)
    sim = rvh.HairSim(cfg)
    with pytest.raises(rvh.RvhError):
        sim.step(DT)                                   # before upload
    st = synth(64, 8, 2.5)
    with pytest.raises(rvh.RvhError):
        sim.upload(st[:32])                            # wrong size
    sim.upload(st)
    with pytest.raises(rvh.RvhError):
        sim.step(DT)                                   # colliders not set
    sim.set_colliders(rvh.scenes.reference_colliders())
    with pytest.raises(rvh.RvhError):
        sim.step(0.0)                                  # dt must be > 0
    sim.step(DT)
    assert sim.kernel_launches() >= 2
    sim.close()


# ---- full-size properties (sizes the oracle does not finish in seconds) ----------------------

def test_full_size_properties_1m_x_32():
    S, N, L = 1 << 20, 32, 2.5
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L)
    rest = np.float32(L) / np.float32(N - 1)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B, rest_length=float(rest))
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    for k in range(3):
        sim.step(DT, 0.1 * k)
    sim.step_phases(DT, 0.5, 1)
    mid = sim.download()
    grid = sim.download_grid()
    sim.close()
    assert np.array_equal(bits(mid[:, 0, 0]), bits(st[:, 0, 0]))                     # roots pinned, indexing exact
    seg = np.linalg.norm(mid[:, 0, 1:, :3] - mid[:, 0, :-1, :3], axis=2)
    assert np.abs(seg / rest - 1).max() <= 2e-5                                       # float32 norm here
    assert np.isfinite(mid).all()
    # density checksum: sum over grid == sum over points of the per-corner truncated weights
    h = np.float32(7.0) / np.float32(64.0)
    g = ((mid[:, 0, 1:, :3] - np.array([-3, -2, -5], np.float32)) / h).reshape(-1, 3)
    f = np.floor(g)
    total = 0
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                cell = f + np.array([a, b, c], np.float32)
                ok = np.all((cell >= 0) & (cell <= 63), axis=1)
                w = np.clip(np.float32(1) - np.abs(g - cell), 0, 1)
                tw = (w[:, 0] * w[:, 1]) * w[:, 2]
                total += int(np.trunc(np.float32(1e6) * tw)[ok].astype(np.int64).sum())
    assert int(grid[:, 3].sum()) == total
    # subset parity on 2048 strands against the oracle (strands only interact through the grid,
    # so the integrate/FTL phase of a subset is independent of the rest)
    sub = np.arange(0, S, S // 2048)[:2048]
    p = orc.default_params(len(sub), N, orc.WIND_B, rest_length=rest)
    # reproduce the state before the last phase-1 call is not available; instead check phase 1 from the initial state
    cfg2 = rvh.default_config(S, N, flags=rvh.WIND_B, rest_length=float(rest))
    sim = rvh.HairSim(cfg2)
    sim.set_colliders(cols)
    sim.upload(st)
    sim.step(DT, 0.25)
    out = sim.download()
    sim.close()
    ref, _ = orc.step(p, cols, DT, 0.25, st[sub])
    check_state(out[sub], ref, float(rest), N, what="1Mx32 subset")


def test_interop_import_rejects_a_bad_handle_without_side_effects():
    """rvh_import_strands_fd needs a VK_KHR_external_memory_fd handle; there is no Vulkan device in this image, so only the
    failure path can be exercised: a bogus fd must come back as an error code + message, and the context must keep working."""
    import ctypes as C
    S, N = 256, 8
    cols = rvh.scenes.reference_colliders()
    st = synth(S, N, 2.5)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON)
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    r = sim.L.rvh_import_strands_fd(sim.ctx, -1, sim.aos_bytes)
    assert r < 0 and sim.L.rvh_last_error(sim.ctx)
    r = sim.L.rvh_import_strands_fd(sim.ctx, 0, sim.aos_bytes - 16)        # smaller than Strand[S]
    assert r == -1
    sim.step(DT, 0.0)
    out = sim.download()
    sim.close()
    rest = np.float32(2.5) / np.float32(N - 1)
    ref, _ = orc.step(orc.default_params(S, N, orc.GRID_ON, rest_length=rest), cols, DT, 0.0, st)
    check_state(out, ref, float(rest), N, what="after failed import")


@pytest.mark.parametrize("S,N,spt", [(160000, 16, 2), (80000, 24, 1)])         # >= 4 x 148 CTAs of 128 threads: the masked kernel is used
def test_collider_candidate_mask_never_drops_a_collision(S, N, spt):
    """In steady-state stepping k_ftl_step runs only the ellipsoid tests its collider candidate mask allows (gather_pack).
    Stepping with a read-back in between forces the unmasked kernel (no gather pending), so the two must agree, on a scene
    whose hair lies on the head, neck and shoulders."""
    L = 2.5
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    p = orc.default_params(S, N, orc.GRID_ON, rest_length=rest)
    st = synth(S, N, L, seed_vel=21)
    for k in range(50):                                   # oracle free run: drape the hair over head, bust and shoulders
        st, _ = orc.step(p, cols, DT, 0.0, st, threads=8)
    inv = cols[1:, 16:32].reshape(-1, 4, 4).transpose(0, 2, 1)                 # Collider::inv, row-major
    pts = np.concatenate([st[:, 0, 1:, :3].reshape(-1, 3), np.ones((S * (N - 1), 1), np.float32)], axis=1)
    inside = (np.linalg.norm(np.einsum("jrc,pc->jpr", inv, pts)[..., :3], axis=-1) <= 1.0)
    assert inside.any(axis=0).mean() > 0.05 and (inside.sum(axis=1) > 0).sum() >= 3, "scene must engage several ellipsoids"
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON, rest_length=float(rest), strands_per_thread=spt)
    a = rvh.HairSim(cfg); a.set_colliders(cols); a.upload(st)
    b = rvh.HairSim(cfg); b.set_colliders(cols); b.upload(st)
    for k in range(4):
        a.step(DT, 0.0)                                    # fused gather + masked collider tests from step 2 on
        b.step(DT, 0.0)
        b.upload(b.download())                             # stand-alone gather, next step tests every ellipsoid
    fa, fb = a.download(), b.download()
    a.close(); b.close()
    assert np.abs(fa[:, 0, :, :3] - fb[:, 0, :, :3]).max() <= 2e-5
    assert np.abs(fa[:, 1, :, :3] - fb[:, 1, :, :3]).max() <= 2e-3


def test_collider_candidate_mask_is_conservative():
    """Every point inside ellipsoid j (|inv_j (p,1)| <= 1, compute.comp:64-67) must lie in a box whose mask byte has bit j."""
    rng = np.random.default_rng(11)
    for cols in (rvh.scenes.bench_colliders(), rvh.scenes.reference_colliders()):
        sim = rvh.HairSim(rvh.default_config(256, 4, flags=rvh.GRID_ON))
        sim.set_colliders(cols)
        mask = sim.collider_mask()
        sim.close()
        assert mask is not None and mask.shape == (32, 32, 32)
        h2 = 2 * 7.0 / 64
        origin = np.array([-3, -2, -5], np.float64)
        for j in range(5):
            X = cols[1 + j, 0:16].reshape(4, 4).T.astype(np.float64)                 # transform: unit sphere -> ellipsoid
            u = rng.normal(size=(200000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
            u *= rng.uniform(0, 1, (200000, 1)) ** (1 / 3)                            # uniform in the unit ball
            p = u @ X[:3, :3].T + X[:3, 3]
            b = np.floor((p - origin) / h2).astype(int)
            ok = np.all((b >= 0) & (b < 32), axis=1)                                  # points outside the grid never consult the mask
            bits = mask[b[ok, 2], b[ok, 1], b[ok, 0]]
            assert np.all(bits & (1 << j)), "ellipsoid %d: %d interior points in unflagged boxes" % (j, int(np.sum((bits & (1 << j)) == 0)))
        # and it prunes: most boxes allow at most two ellipsoids
        assert (np.unpackbits(mask.reshape(-1, 1), axis=1).sum(axis=1) <= 2).mean() > 0.8
    sim = rvh.HairSim(rvh.default_config(256, 4, flags=0))                            # grid off: no mask
    sim.set_colliders(rvh.scenes.bench_colliders())
    assert sim.collider_mask() is None
    sim.close()
