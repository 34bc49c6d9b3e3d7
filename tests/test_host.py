"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/rvh.h
declares, host maths agree with the oracle, and the product path fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc
import rvh_b200 as rvh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    L = rvh.load_library()
    header = open(os.path.join(ROOT, "include", "rvh.h")).read()
    declared = set(re.findall(r"\b(rvh_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(rvh.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.rvh_abi_version() == 3


def test_struct_sizes_match_reference_layouts():
    # Strand = 48*N bytes, Collider = 192, GridCell = 16, StrandDrawIndirect = 16 (Strand.h, Scene.h)
    assert C.sizeof(rvh.RvhConfig) == 4 * 19
    cfg = rvh.default_config(900, 10)
    assert cfg.num_strands * 48 * cfg.num_points == 432000          # Renderer.cpp:967 descriptor range
    assert abs(cfg.rest_length - np.float32(2.5) / np.float32(9.0)) == 0
    assert (cfg.grid_dim, cfg.grid_extent, cfg.grid_scale, cfg.friction) == (64, 7.0, 1e6, np.float32(0.08))
    assert cfg.flags == rvh.GRID_ON


def test_oracle_and_product_configs_agree():
    p = orc.default_params(1234, 17)
    c = rvh.default_config(1234, 17)
    for f in ("rest_length", "gravity_y", "damping", "vmax", "penalty_k", "sphere_radius", "grid_dim",
              "grid_extent", "grid_scale", "friction"):
        assert getattr(p, f) == getattr(c, f), f
    assert list(p.grid_origin) == list(c.grid_origin)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_gpu():
    cfg = rvh.default_config(64, 8)
    with pytest.raises(rvh.RvhError) as e:
        rvh.HairSim(cfg)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_arguments():
    L = rvh.load_library()
    ctx = C.c_void_p()
    cfg = rvh.default_config(0, 10)
    assert L.rvh_create(C.byref(ctx), C.byref(cfg)) == -1
    cfg = rvh.default_config(10, 1)
    assert L.rvh_create(C.byref(ctx), C.byref(cfg)) == -1
    assert L.rvh_create(None, C.byref(cfg)) == -1
    assert b"num_points" in L.rvh_last_error(None) or b"null" in L.rvh_last_error(None)


def test_host_collider_maths_vs_oracle(golden_c1):
    mine = rvh.scenes.reference_colliders()
    ref = golden_c1["colliders"]
    assert np.abs(mine - ref).max() <= 1e-6
    moved = rvh.collider_translate(mine[0], golden_c1["sphere_translation"])
    assert np.abs(moved - golden_c1["sphere_moved"]).max() <= 1e-6


def test_host_wind_fbm_vs_oracle():
    for t in (0.0, 0.5, 1.2345, 7.0, 100.25):
        assert abs(rvh.wind_fbm(t) - orc.fbm_time(t)) <= 2e-6


def test_synthetic_head_is_shardable_and_at_rest_spacing():
    full = rvh.scenes.synthetic_head(1000, 16, 0.4)
    lo, hi = rvh.scenes.shard_range(1000, 1, 3)
    part = rvh.scenes.synthetic_head(hi - lo, 16, 0.4, first_strand=lo)
    assert np.array_equal(full[lo:hi], part)
    seg = np.linalg.norm(full[:, 0, 1:, :3] - full[:, 0, :-1, :3], axis=2)
    assert np.abs(seg / (0.4 / 15) - 1).max() < 1e-4
    assert np.all(full[:, 0, :, 3] == 1) and np.all(full[:, 1, :, 2] == -1)
    # roots lie on the head ellipsoid (collider 1): |inv * root| == 1
    cols = rvh.scenes.reference_colliders()
    inv = cols[1, 16:32].reshape(4, 4).T
    q = full[:, 0, 0, :3] @ inv[:3, :3].T + inv[:3, 3]
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-4


def test_shard_ranges_partition_all_strands():
    for S in (1, 7, 900, 1000003):
        for R in (1, 2, 3, 8):
            r = [rvh.scenes.shard_range(S, k, R) for k in range(R)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[i][1] == r[i + 1][0] for i in range(R - 1))


def _c_abi_smoke(tmp_path):
    """include/rvh.h compiles as C11 and links against librvh.so; without a GPU the program must stop at rvh_create with the
    no-fallback error (exit 3), with one it runs ten steps (exit 0)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(rvh.library_path())
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_smoke.c"),
                           "-L", libdir, "-lrvh", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "abi 3" in out.stdout and "config 76 bytes" in out.stdout
    if _has_gpu():
        assert out.returncode == 0 and "10 steps ok" in out.stdout and "{900,1,0,0}" in out.stdout, out.stdout
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stdout, out.stdout


def test_c_abi_is_usable_from_plain_c(tmp_path):
    _c_abi_smoke(tmp_path)


@pytest.mark.gpu
def test_c_abi_smoke_runs_ten_steps_on_the_gpu(tmp_path):
    assert _has_gpu()
    _c_abi_smoke(tmp_path)


def test_host_collider_maths_random_transforms_vs_oracle():
    """Scene.h:28-38 for arbitrary translation / rotation (degrees) / scale: the library's own matrix code against the oracle's
    glm restatement (which is pinned bit-exactly to the reference's glm build)."""
    from hypothesis import given, settings, strategies as st
    f = lambda lo, hi: st.floats(lo, hi, allow_nan=False, width=32)

    @settings(max_examples=150, deadline=None)
    @given(st.tuples(f(-5, 5), f(-5, 5), f(-5, 5)), st.tuples(f(-180, 180), f(-180, 180), f(-180, 180)), st.tuples(f(0.25, 3), f(0.25, 3), f(0.25, 3)),
           st.tuples(f(-2, 2), f(-2, 2), f(-2, 2)))
    def check(t, r, s, move):
        mine, ref = rvh.collider_build(t, r, s), orc.collider_build(t, r, s)
        scale = np.abs(ref).max()
        assert np.abs(mine - ref).max() <= 2e-5 * max(1.0, scale)              # inverse of a 0.25-scaled matrix has entries up to 4
        # transform * inv == identity, invTrans == inv^T
        X, I, IT = (mine[16 * k:16 * k + 16].reshape(4, 4).T.astype(np.float64) for k in range(3))
        assert np.abs(X @ I - np.eye(4)).max() <= 1e-4
        assert np.abs(IT - I.T).max() <= 1e-6
        a, b = rvh.collider_translate(mine, move), orc.collider_translate(ref, move)
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max())
    check()
