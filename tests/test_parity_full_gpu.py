"""Full-size GPU parity (VERDICT r01 "parity holes"): the kernel instantiation the benchmark times, the BASELINE.json
configurations at their own sizes with the grid on, and a census of the collider decisions.

All through the C ABI, against the CPU oracle (`orc.step(..., threads=)` = orc_step_parallel, ~1-2 s per step at these sizes).
Tolerances are the north star's: max |position error| <= 1e-4 * strand length per step, integer grid bit-exact given the
same inputs.
"""
import os

import numpy as np
import pytest

import orc
import rvh_b200 as rvh
from test_parity_gpu import DT, POS_TOL_REL, bits, check_state

pytestmark = pytest.mark.gpu
THREADS = min(32, os.cpu_count() or 8)


def draped(S, N, L, oflags, steps, T0=0.0):
    """Oracle free run from the synthetic head: hair on the head / shoulders, busy grid, wind-driven velocities."""
    cols = rvh.scenes.bench_colliders()
    rest = np.float32(L) / np.float32(N - 1)
    p = orc.default_params(S, N, oflags, rest_length=rest)
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    for k in range(steps):
        st, _ = orc.step(p, cols, DT, np.float32(T0) + np.float32(k) * DT, st, threads=THREADS)
    return st, cols, p, rest


def test_two_step_protocol_runs_the_benchmark_kernel_against_the_oracle():
    """bench.py's steady-state step is k_ftl_step<2, wind, 105, gather> -- packed pairs, wind B, collider candidate mask, the
    previous grid's gather fused into the load -- which needs a pending gather, i.e. a SECOND step without a read-back in
    between, and >= 4 x 148 CTAs.  Upload oracle state k-1, run two GPU steps, compare with oracle state k+1.

    Tolerance: the per-step bar (1e-4 * L) times 2, the stated two-step factor: step k+1 starts from a GPU state that is within
    the one-step error of the oracle's (measured ~4e-7 * L), and SURVEY.md section 7 measured free-running amplification of
    ~1.5x per step."""
    S, N, L = 655360, 16, 2.5                                  # 2,560 CTAs of 256 strands: masked kernel + fused clear are used
    st, cols, p, rest = draped(S, N, L, orc.GRID_ON | orc.WIND_B, steps=3, T0=0.2)
    t0 = np.float32(0.2) + np.float32(3) * DT
    ref1, _ = orc.step(p, cols, DT, t0, st, threads=THREADS)
    ref2, ref_grid2 = orc.step(p, cols, DT, t0 + DT, ref1, threads=THREADS)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B, rest_length=float(rest), strands_per_thread=2)
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    assert sim.collider_mask() is not None
    sim.step(DT, float(t0))                                    # k_ftl_step<2,1,5,0>, splat, finalize; gather left pending
    sim.step(DT, float(t0 + DT))                               # k_ftl_step<2,1,105,1>: the benchmark's kernel
    out = sim.download()                                       # applies the second step's gather stand-alone
    grid = sim.download_grid()
    sim.close()
    Ltot = float(rest) * (N - 1)
    perr = np.abs(out[:, 0, :, :3] - ref2[:, 0, :, :3]).max()
    verr = np.abs(out[:, 1, :, :3] - ref2[:, 1, :, :3]).max()
    print("two-step: pos err %.3e (%.2e L), vel err %.3e" % (perr, perr / Ltot, verr))
    assert perr <= 2 * POS_TOL_REL * Ltot
    assert verr <= 2 * POS_TOL_REL * Ltot / float(DT)
    assert np.array_equal(bits(out[:, 0, 0]), bits(st[:, 0, 0]))
    seg = np.linalg.norm(out[:, 0, 1:, :3].astype(np.float64) - out[:, 0, :-1, :3], axis=2)
    assert np.abs(seg / float(rest) - 1).max() <= 1e-5
    # the second step's grid: same occupied cells up to boundary flips, totals within one unit per contribution
    assert abs(int(grid[:, 3].sum()) - int(ref_grid2[:, 3].sum())) <= 8 * S * N


@pytest.mark.parametrize("name,S,N,L", [("C3", 100000, 64, 2.5), ("C4", 1000000, 16, 0.4)])
def test_baseline_configs_full_size_one_step_with_grid(name, S, N, L):
    """BASELINE.json configs[2] and configs[3] at their own sizes, grid on: one step from a settled oracle state against
    orc_step_parallel; the integer grid bit-exact against the oracle's splat of the downloaded mid-state."""
    st, cols, p, rest = draped(S, N, L, orc.GRID_ON, steps=2)
    T = np.float32(2) * DT
    ref, ref_grid = orc.step(p, cols, DT, T, st, threads=THREADS)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON, rest_length=float(rest))
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    sim.step_phases(DT, float(T), 1)                           # integrate + FTL + corrected velocity + splat
    mid = sim.download()
    grid = sim.download_grid()
    sim.step_phases(DT, float(T), 2)                           # finalize + gather
    out = sim.download()
    sim.close()
    perr, verr = check_state(out, ref, float(rest), N, what=name)
    print("%s full size: pos err %.3e vel err %.3e" % (name, perr, verr))
    _, want_grid = orc.phase_splat(p, DT, mid)                 # zero correctionVecs in `mid`: a pure splat of the GPU's own mid-state
    assert np.array_equal(grid, want_grid), "%s grid integers differ in %d cells" % (name, int(np.any(grid != want_grid, axis=1).sum()))
    assert abs(int(grid[:, 3].sum()) - int(ref_grid[:, 3].sum())) <= 8 * S * N
    assert grid[:, 3].max() > 2 ** 31 or name == "C3"          # C4's fur density overflows an int32 cell: int64 matters


def test_collider_decision_census_on_a_draped_scene():
    """rsqrt.approx / FMA contraction / squared-distance compares can flip `inside collider j` for a point within an ulp of the
    surface.  The force is continuous there, but the division by the hit count (compute.comp:182-184) is not.  Count the
    flips on a scene lying on head, neck, bust and shoulders: < 1e-5 of the points."""
    S, N, L = 200000, 32, 2.5
    st, cols, p, rest = draped(S, N, L, orc.GRID_ON, steps=40)
    want = orc.hit_masks(p, cols, st)
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON, rest_length=float(rest))
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    got = sim.hit_masks()
    sim.step(DT, 0.0)
    out = sim.download()
    sim.close()
    inside = (want != 0).mean()
    multi = (np.unpackbits(want.reshape(-1, 1), axis=1).sum(axis=1) >= 2).mean()
    flips = int((got != want).sum())
    print("census: %.1f%% of points inside a collider, %.2f%% inside two or more, %d decision flips of %d points" % (100 * inside, 100 * multi, flips, S * (N - 1)))
    assert inside > 0.05, "scene must engage the colliders"
    assert flips <= 1e-5 * S * (N - 1)
    # (the other data-dependent branches -- the velocity clamp at |v| = vmax, compute.comp:198-200, and the sphere / ellipsoid
    # penalty at zero depth -- are continuous across their thresholds, so a flipped decision there changes nothing beyond rounding)
    ref, _ = orc.step(p, cols, DT, 0.0, st, threads=THREADS)
    check_state(out, ref, float(rest), N, what="census scene")


def test_free_running_ten_steps_stay_within_the_chaos_bound(golden_c1):
    """Free-running comparison is only meaningful for <~10 steps (two CPU builds of the same text differ by
    2.3e-5 * L after 10 steps, 6.4e-2 * L after 100).  From the reference's own violent initial state (Strand.cpp:157-175)."""
    st, cols = golden_c1["state0"], golden_c1["colliders"]
    sim = rvh.HairSim(rvh.default_config(900, 10, flags=rvh.GRID_ON | rvh.GRID_INT32_WRAP))
    sim.set_colliders(cols)
    sim.upload(st)
    p = orc.default_params(900, 10, orc.GRID_ON | orc.GRID_INT32_WRAP)
    ref = st.copy()
    errs = []
    for k in range(10):
        sim.step(DT, float(k) * float(DT))
        ref, _ = orc.step(p, cols, DT, np.float32(k) * DT, ref)
        errs.append(float(np.abs(sim.download()[:, 0, :, :3] - ref[:, 0, :, :3]).max()) / 2.5)
    sim.close()
    print("free run, max |dp| / L per step:", " ".join("%.1e" % e for e in errs))
    assert errs[0] <= POS_TOL_REL
    assert errs[-1] <= 5e-5, "10 free-running steps drifted %.2e L (measured 1.1e-5 on B200)" % errs[-1]


# ---- the fast paths of rvh_step_n / rvh_step_host must be invisible in the results ----------------------------------

def _fresh(S, N, L, flags, st, cols, spt=0):
    rest = float(np.float32(L) / np.float32(N - 1))
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=rest, strands_per_thread=spt))
    sim.set_colliders(cols)
    sim.upload(st)
    return sim


def _plain(monkeypatch, S, N, L, flags, st, cols, spt=0):
    """A context whose steps never take the persistent small-scene kernel (k_scene_step): the launch-per-kernel path."""
    monkeypatch.setenv("RVH_SCENE_CTAS", "0")                   # read by rvh_create
    sim = _fresh(S, N, L, flags, st, cols, spt)
    monkeypatch.delenv("RVH_SCENE_CTAS")
    return sim


@pytest.mark.parametrize("S,N,flags,spt", [(16384, 32, rvh.WIND_B, 0), (16384, 32, rvh.WIND_B, 2), (3000, 16, rvh.WIND_A, 1), (900, 10, 0, 0)])
def test_step_n_without_grid_runs_many_steps_per_launch_with_the_same_result(S, N, flags, spt):
    """Grid off: rvh_step_n puts up to 32 steps into ONE launch (k_ftl_step<..., MULTI>, per-step wind scalars from a host
    table).  40 steps that way == 40 calls of rvh_step, bit for bit (C2's shape first)."""
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    a = _fresh(S, N, 2.5, flags, st, cols, spt)
    la = a.kernel_launches()
    a.step_n(40, DT, 0.25)
    assert a.kernel_launches() - la == 2                        # 32 + 8 steps
    fa = a.download()
    a.close()
    # same float sequence of step times as the library (t += dt in float32): 40 free-running steps amplify a last-bit difference
    c = _fresh(S, N, 2.5, flags, st, cols, spt)
    t = np.float32(0.25)
    for k in range(40):
        c.step(DT, float(t))
        t = np.float32(t + np.float32(DT))
    fc = c.download()
    c.close()
    assert np.array_equal(bits(fa), bits(fc)), "multi-step launch differs from single steps"


def test_step_n_replays_a_cuda_graph_on_small_grid_scenes_with_the_same_result(golden_c1, monkeypatch):
    """Grid on, wind off, small scene (the shipped C1 scene) with the persistent kernel switched off: rvh_step_n captures one
    steady-state step as a CUDA graph and replays it.  Same kernels, same order: bit-identical to calling rvh_step."""
    st, cols = golden_c1["state0"], golden_c1["colliders"]
    flags = rvh.GRID_ON | rvh.GRID_INT32_WRAP
    monkeypatch.setenv("RVH_SCENE_CTAS", "0")
    a = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); a.set_colliders(cols); a.upload(st)
    b = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); b.set_colliders(cols); b.upload(st)
    a.step_n(25, DT, 0.0)
    a.step_n(5, DT, 25 * float(DT))                             # reuses the instantiated graph
    for k in range(30):
        b.step(DT, k * float(DT))
    fa, fb, ga, gb = a.download(), b.download(), a.download_grid(), b.download_grid()
    a.set_colliders(cols)                                       # invalidates the captured step (its kernel parameters hold the colliders)
    a.step_n(6, DT, 0.0)
    for k in range(6):
        b.step(DT, 0.0)
    fa2, fb2 = a.download(), b.download()
    a.close(); b.close()
    assert np.array_equal(bits(fa), bits(fb)) and np.array_equal(ga, gb)
    assert np.array_equal(bits(fa2), bits(fb2))


@pytest.mark.parametrize("S,N,L,flags", [(150000, 8, 0.4, rvh.GRID_ON), (140001, 12, 2.5, rvh.GRID_ON | rvh.WIND_B), (131072, 10, 2.5, rvh.WIND_B)])
def test_pipelined_step_host_equals_upload_step_download(S, N, L, flags):
    """Above 128K strands rvh_step_host pipelines chunked copies against the kernels and returns positions before the grid
    is complete.  The host buffer must end bit-identical to upload + step + download (which Morton-sorts the strands)."""
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    rng = np.random.default_rng(5)
    st[:, 1, 1:, :3] += rng.normal(scale=0.3, size=(S, N - 1, 3)).astype(np.float32)
    ref_sim = _fresh(S, N, L, flags, st, cols)
    ref_sim.step(DT, 0.4)
    ref = ref_sim.download()
    ref_sim.close()
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=float(np.float32(L) / np.float32(N - 1))))
    sim.set_colliders(cols)
    buf = st.copy()
    buf[:, 2] = 123.0
    sim.step_host(buf, DT, 0.4)
    again = sim.download()                                      # the device state after the pipelined call, through the normal path
    sim.step(DT, 0.4 + float(DT))                               # and the context keeps stepping from it
    sim.close()
    assert np.array_equal(bits(buf[:, 0:2]), bits(ref[:, 0:2]))
    assert np.all(buf[:, 2] == 123.0)
    assert np.array_equal(bits(again[:, 0:2]), bits(ref[:, 0:2]))


def test_interop_pack_and_indirect_args_land_in_external_device_buffers():
    """Everything behind rvh_import_strands_fd / rvh_import_indirect_fd except the import itself (no Vulkan device here): with
    caller-owned device buffers standing in for the imported VkBuffers, every rvh_step must leave the reference's vertex-buffer
    layout Strand[S] (Strand.h:11-15, curvePoints first) and StrandDrawIndirect {S,1,0,0} in them -- the same bytes
    rvh_download_strands_aos returns, gather applied."""
    import torch
    S, N, L = 5000, 10, 2.5
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    ext = torch.full((S * 3 * N * 4,), 7.0, dtype=torch.float32, device="cuda")
    ind = torch.full((4,), 99, dtype=torch.int32, device="cuda")
    sim = rvh.HairSim(rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B | rvh.KEEP_CORRECTION))
    sim.set_colliders(cols)
    sim.upload(st)
    assert sim.L.rvh_debug_set_interop_device_buffers(sim.ctx, ext.data_ptr(), ext.numel() * 4, ind.data_ptr()) == 0
    for k in range(3):
        sim.step(DT, 0.1 * k)
        sim.sync()
        got = ext.cpu().numpy().reshape(S, 3, N, 4)
        want = sim.download()
        assert np.array_equal(bits(got), bits(want)), "step %d: external Strand[S] differs from the download" % k
        assert ind.cpu().tolist() == [S, 1, 0, 0]
    assert sim.L.rvh_debug_set_interop_device_buffers(sim.ctx, None, 0, None) == 0
    sim.step(DT, 0.3)
    sim.sync()
    assert np.array_equal(bits(ext.cpu().numpy().reshape(S, 3, N, 4)), bits(want))      # detached: untouched
    # the import entry points themselves: bogus handles are errors, not crashes, and leave the context usable
    assert sim.L.rvh_import_indirect_fd(sim.ctx, -1, 16) < 0 and sim.L.rvh_last_error(sim.ctx)
    assert sim.L.rvh_import_indirect_fd(sim.ctx, 0, 8) == -1
    assert sim.L.rvh_import_semaphore_fd(sim.ctx, -1) < 0
    sim.step(DT, 0.4)
    sim.close()


def test_resident_steps_after_a_pipelined_host_step_run_on_morton_order_again():
    """The chunk pipeline leaves the strands in the caller's order; the next resident step must restore the Morton order (the
    splat's warp aggregation depends on it: ~10x slower otherwise) without changing a bit of the result."""
    S, N, L = 200000, 16, 2.5
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    flags = rvh.GRID_ON | rvh.WIND_B
    a = rvh.HairSim(rvh.default_config(S, N, flags=flags)); a.set_colliders(cols)
    buf = st.copy()
    a.step_host(buf, DT, 0.1)
    a.step(DT, 0.15)                                            # the first resident step re-sorts (not timed)
    ms_a = a.step_n(20, DT, 0.2) / 20
    fa = a.download()
    a.close()
    b = _fresh(S, N, L, flags, st, cols)
    b.step(DT, 0.1)
    b.upload(b.download())
    b.step(DT, 0.15)
    ms_b = b.step_n(20, DT, 0.2) / 20
    fb = b.download()
    b.close()
    assert np.array_equal(bits(fa[:, 0:2]), bits(fb[:, 0:2]))
    # on unsorted strands the splat alone is ~10x slower: a factor of two separates the cases with room for timing noise
    assert ms_a <= 2.0 * ms_b + 0.05, "resident steps after rvh_step_host are slow: %.3f ms vs %.3f ms" % (ms_a, ms_b)


@pytest.mark.parametrize("n", [1, 2, 3, 33, 65])
def test_step_n_batch_edges_with_correction_vectors(n):
    """Batch boundaries of the many-steps launch (32 per launch) and correctionVecs (written by every step, RVH_KEEP_CORRECTION)."""
    S, N = 2000, 12
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    flags = rvh.WIND_B | rvh.KEEP_CORRECTION
    a, b = _fresh(S, N, 2.5, flags, st, cols), _fresh(S, N, 2.5, flags, st, cols)
    a.step_n(n, DT, 0.5)
    t = np.float32(0.5)
    for k in range(n):
        b.step(DT, float(t))
        t = np.float32(t + np.float32(DT))
    fa, fb = a.download(), b.download()
    a.close(); b.close()
    assert np.array_equal(bits(fa), bits(fb))
    assert np.abs(fa[:, 2, 1:, :3]).max() > 0


@pytest.mark.parametrize("flags", [rvh.GRID_ON | rvh.REPULSION_ON, rvh.GRID_ON | rvh.KEEP_CORRECTION, rvh.GRID_ON | rvh.GRID_INT32_WRAP])
def test_graph_replay_with_extension_and_layout_flags(flags, monkeypatch):
    S, N = 6000, 16
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    monkeypatch.setenv("RVH_SCENE_CTAS", "0")                   # the graph path (k_scene_step would take the last two flag sets)
    a, b = _fresh(S, N, 2.5, flags, st, cols), _fresh(S, N, 2.5, flags, st, cols)
    a.step_n(9, DT, 0.0)
    for k in range(9):
        b.step(DT, 0.0)
    fa, fb, ga, gb = a.download(), b.download(), a.download_grid(), b.download_grid()
    a.close(); b.close()
    assert np.array_equal(bits(fa), bits(fb)) and np.array_equal(ga, gb)


def test_pipelined_step_host_uneven_chunks_and_extensions(monkeypatch):
    """Three uneven chunks (RVH_HOST_CHUNKS) with the head SDF and repulsion on: same bytes as upload + step + download."""
    monkeypatch.setenv("RVH_HOST_CHUNKS", "3")
    S, N, L = 131072 + 300, 8, 2.5
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    flags = rvh.GRID_ON | rvh.SDF_ON | rvh.REPULSION_ON | rvh.WIND_A
    dim, origin, cell = [41, 63, 35], np.array([-2.0, -2.2, -1.8], np.float32), 0.1
    rest = float(np.float32(L) / np.float32(N - 1))
    ref_sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=rest))
    ref_sim.set_colliders(cols); ref_sim.bake_head_sdf_from_colliders(dim, origin, cell)
    ref_sim.upload(st); ref_sim.step(DT, 0.2)
    ref = ref_sim.download(); ref_sim.close()
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=rest))
    sim.set_colliders(cols); sim.bake_head_sdf_from_colliders(dim, origin, cell)
    buf = st.copy()
    sim.step_host(buf, DT, 0.2)
    sim.close()
    assert np.array_equal(bits(buf[:, 0:2]), bits(ref[:, 0:2]))


# ---- small scenes: whole steps inside one persistent cooperative launch (k_scene_step) -------------------------------------

@pytest.mark.parametrize("S,N,flags,spt", [(900, 10, rvh.GRID_ON | rvh.GRID_INT32_WRAP, 0), (3000, 16, rvh.GRID_ON | rvh.WIND_A, 1),
                                           (6000, 16, rvh.GRID_ON | rvh.WIND_B | rvh.KEEP_CORRECTION, 2), (9000, 32, rvh.GRID_ON | rvh.WIND_B, 0),
                                           (18000, 8, rvh.GRID_ON, 2), (130, 2, rvh.GRID_ON | rvh.WIND_B, 0), (1, 5, rvh.GRID_ON, 0),
                                           (257, 3, rvh.GRID_ON | rvh.GRID_INT32_WRAP | rvh.WIND_A, 2)])
def test_scene_step_kernel_equals_launch_per_kernel_steps(S, N, flags, spt, monkeypatch):
    """rvh_step_n on a small grid scene: the steps run inside k_scene_step, up to 32 per launch -- FTL without gather | grid barrier |
    splat | grid barrier | fully parallel gather straight from the int64 accumulators | grid barrier, the grid clear riding on the
    CTAs the FTL phase leaves idle, per-step wind scalars from the host table.  40 steps == 40 launch-per-kernel steps (fused
    gather through the finalized float grid), bit for bit, state and integer grid."""
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    a = _fresh(S, N, 2.5, flags, st, cols, spt)
    la = a.kernel_launches()
    a.step_n(40, DT, 0.25)
    used = a.kernel_launches() - la
    fa, ga = a.download(), a.download_grid()
    a.close()
    b = _plain(monkeypatch, S, N, 2.5, flags, st, cols, spt)
    t = np.float32(0.25)
    lb = b.kernel_launches()
    for k in range(40):
        b.step(DT, float(t))
        t = np.float32(t + np.float32(DT))
    plain_used = b.kernel_launches() - lb
    fb, gb = b.download(), b.download_grid()
    b.close()
    assert used == 2 and plain_used >= 3 * 40, (used, plain_used)    # 32 + 8 steps
    assert np.array_equal(bits(fa), bits(fb)), "persistent small-scene kernel differs from launch-per-kernel steps"
    assert np.array_equal(ga, gb)
    assert np.abs(ga).max() > 0


def test_rvh_step_takes_one_launch_per_step_on_a_small_scene(golden_c1, monkeypatch):
    """The per-frame call itself (rvh_step, as Renderer::Frame would issue it) on the shipped C1 scene:
    one k_scene_step launch per step, same state and grid as the launch-per-kernel path; colliders moving between steps
    (Scene::translateSphere) and a read-back in the middle do not disturb it."""
    st, cols = golden_c1["state0"], np.array(golden_c1["colliders"], copy=True)
    flags = rvh.GRID_ON | rvh.GRID_INT32_WRAP
    a = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); a.set_colliders(cols); a.upload(st)
    monkeypatch.setenv("RVH_SCENE_CTAS", "0")
    b = rvh.HairSim(rvh.default_config(900, 10, flags=flags)); b.set_colliders(cols); b.upload(st)
    monkeypatch.delenv("RVH_SCENE_CTAS")
    a.step(DT, 0.0); b.step(DT, 0.0)
    per_step = []
    for k in range(1, 12):
        if k == 4:
            cols2 = cols.copy(); cols2[0, 12:15] += np.float32(0.05)     # the sphere's translation column (Scene.cpp:110-136)
            a.set_colliders(cols2); b.set_colliders(cols2)
        if k == 7:
            assert np.array_equal(bits(a.download()), bits(b.download()))   # applies the pending gather on both sides
        l0 = a.kernel_launches()
        a.step(DT, k * float(DT)); b.step(DT, k * float(DT))
        per_step.append(a.kernel_launches() - l0)
    fa, fb, ga, gb = a.download(), b.download(), a.download_grid(), b.download_grid()
    a.close(); b.close()
    assert np.array_equal(bits(fa), bits(fb)) and np.array_equal(ga, gb)
    assert per_step == [1] * 11, per_step


# ---- grid off, small scenes: the wavefront over the steps (k_ftl_wave) ------------------------------------------------------

@pytest.mark.parametrize("S,N,n,flags", [(16384, 5, 7, rvh.WIND_B), (16384, 8, 6, rvh.WIND_A | rvh.KEEP_CORRECTION), (16000, 32, 5, 0),
                                         (1000, 3, 9, rvh.WIND_B), (1000, 4, 17, rvh.WIND_B | rvh.KEEP_CORRECTION), (27000, 16, 40, rvh.WIND_B),
                                         (300, 2, 5, rvh.WIND_B)])
def test_step_wavefront_partial_batches_short_strands_same_bits(S, N, n, flags, monkeypatch):
    """k_ftl_wave: W = 4 or 8 lanes share a strand, lane j runs step s0 + j two rows behind lane j - 1.  Partial batches (n not a
    multiple of W), the shortest strands (N = 3: one interior row; N = 2 falls back to the many-steps kernel), both lane counts
    (16K strands -> 4 lanes, 1K -> 8): bit-identical to single steps AND to the many-steps kernel k_ftl_step<MULTI>."""
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, 2.5, colliders=cols)
    a = _fresh(S, N, 2.5, flags, st, cols)
    la = a.kernel_launches()
    a.step_n(n, DT, 0.5)
    assert a.kernel_launches() - la == (n + 31) // 32
    fa = a.download(); a.close()
    monkeypatch.setenv("RVH_WAVE_STEPS", "0")
    m = _fresh(S, N, 2.5, flags, st, cols)
    monkeypatch.delenv("RVH_WAVE_STEPS")
    m.step_n(n, DT, 0.5)
    fm = m.download(); m.close()
    b = _fresh(S, N, 2.5, flags, st, cols)
    t = np.float32(0.5)
    for k in range(n):
        b.step(DT, float(t))
        t = np.float32(t + np.float32(DT))
    fb = b.download(); b.close()
    assert np.array_equal(bits(fa), bits(fb)), "step wavefront differs from single steps"
    assert np.array_equal(bits(fm), bits(fb)), "many-steps kernel differs from single steps"


def test_step_wavefront_against_the_oracle():
    """The kernel the C2 bench line times (k_ftl_wave, 4 lanes per strand at 16K x 32), compared with the oracle directly: one
    launch = 4 steps in flight at once, against 4 free-running oracle steps (chaos bound of SURVEY.md section 7, as the ten-step test)."""
    S, N, L = 16384, 32, 2.5
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    sim = _fresh(S, N, L, rvh.WIND_B, st, cols)
    l0 = sim.kernel_launches()
    sim.step_n(4, DT, 0.5)
    assert sim.kernel_launches() - l0 == 1
    got = sim.download(); sim.close()
    p = orc.default_params(S, N, orc.WIND_B, rest_length=np.float32(L) / np.float32(N - 1))
    ref, t = st.copy(), np.float32(0.5)
    for k in range(4):
        ref, _ = orc.step(p, cols, DT, t, ref)
        t = np.float32(t + np.float32(DT))
    err = float(np.abs(got[:, 0, :, :3] - ref[:, 0, :, :3]).max()) / L
    print("wavefront, 4 steps, max |dp| / L = %.2e" % err)
    assert err <= 2e-5
    assert np.array_equal(bits(got[:, 0, 0]), bits(ref[:, 0, 0]))
