"""Small run through every kernel of librvh.so, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvh_b200 as rvh

DT = float(np.float32(1 / 60))
cols = rvh.scenes.bench_colliders()
origin, cell, dim = np.array([-2.0, -2.2, -1.8], np.float32), 0.1, [41, 63, 35]
m = np.load(os.path.join(ROOT, "tests", "golden", "mannequin_head_mesh.npz"))
for S, N, spt in ((2048, 12, 1), (1500, 10, 2)):
    rest = float(np.float32(2.5) / np.float32(N - 1))
    for flags in (rvh.GRID_ON | rvh.WIND_B, rvh.GRID_ON | rvh.SDF_ON | rvh.REPULSION_ON, rvh.GRID_ON | rvh.SDF_ON | rvh.SDF_TMA | rvh.WIND_A, rvh.KEEP_CORRECTION):
        sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=rest, strands_per_thread=spt))
        sim.set_colliders(cols)
        if flags & rvh.SDF_ON:
            if flags & rvh.SDF_TMA:
                sim.bake_head_sdf_from_colliders(dim, origin, cell)
            else:
                sim.bake_head_sdf_from_mesh(m["v"] * np.float32(0.98), m["tri"][:600], dim, origin, cell)
        sim.upload(rvh.scenes.synthetic_head(S, N, 2.5))
        for k in range(3):
            sim.step(DT, 0.1 * k)
        out = sim.download()
        sim.step_phases(DT, 0.0, 1); sim.step_phases(DT, 0.0, 2)
        g = sim.download_grid()
        pw, tu, ms = sim.expand(5, 9)
        sim.init_synthetic_head(0, 2.5, 8)
        sim.step(DT, 0.0)
        assert np.isfinite(sim.download()).all() and np.isfinite(pw).all()
        sim.close()
        print("ok", S, N, spt, flags, flush=True)

# ---- round 2: the new paths -------------------------------------------------------------------------------------------
import torch  # noqa: E402  (device buffers standing in for imported Vulkan memory)
# many steps per launch (grid off), CUDA-graph replay (grid on, wind off), hit-mask hook
for S, N, flags, spt in ((3000, 10, rvh.WIND_B, 1), (3000, 10, rvh.WIND_A, 2), (900, 10, rvh.GRID_ON, 0)):
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, strands_per_thread=spt))
    sim.set_colliders(cols)
    sim.upload(rvh.scenes.synthetic_head(S, N, 2.5))
    sim.step_n(40, DT, 0.0)
    sim.step_n(7, DT, 1.0)
    hm = sim.hit_masks()
    assert np.isfinite(sim.download()).all() and hm.shape == (S, N)
    sim.close()
    print("ok step_n", S, N, flags, flush=True)
# pipelined rvh_step_host (>= 128K strands), one-cell / two-cell / multi-pass splat rows, wide fused clear
S, N = 131072 + 77, 6
for flags in (rvh.GRID_ON | rvh.WIND_B, rvh.WIND_B, rvh.GRID_ON | rvh.REPULSION_ON):
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, rest_length=float(np.float32(0.4) / np.float32(N - 1))))
    sim.set_colliders(cols)
    buf = rvh.scenes.synthetic_head(S, N, 0.4)
    sim.step_host(buf, DT, 0.2)
    sim.step(DT, 0.3); sim.step(DT, 0.4)
    assert np.isfinite(buf).all() and np.isfinite(sim.download()).all()
    sim.close()
    print("ok step_host", flags, flush=True)
# per-step pack into external device buffers + indirect args
S, N = 2500, 10
ext = torch.zeros(S * 3 * N * 4, dtype=torch.float32, device="cuda"); ind = torch.zeros(4, dtype=torch.int32, device="cuda")
sim = rvh.HairSim(rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.KEEP_CORRECTION))
sim.set_colliders(cols)
sim.upload(rvh.scenes.synthetic_head(S, N, 2.5))
assert sim.L.rvh_debug_set_interop_device_buffers(sim.ctx, ext.data_ptr(), ext.numel() * 4, ind.data_ptr()) == 0
sim.step(DT, 0.0); sim.step(DT, 0.1); sim.sync()
assert ind.cpu().tolist() == [S, 1, 0, 0]
sim.close()
print("ok interop hook", flush=True)

# ---- late round 2: persistent small-scene kernel (k_scene_step), sparse-row splat, step wavefront (k_ftl_wave), graph replay kept reachable
for S, N, flags, spt, env in ((900, 10, rvh.GRID_ON | rvh.GRID_INT32_WRAP, 0, {}), (3000, 12, rvh.GRID_ON | rvh.WIND_B | rvh.KEEP_CORRECTION, 2, {}),
                              (900, 10, rvh.GRID_ON, 0, {"RVH_SCENE_CTAS": "0"}), (16384, 6, rvh.WIND_B, 1, {}), (1000, 3, rvh.WIND_A | rvh.KEEP_CORRECTION, 0, {})):
    os.environ.update(env)
    sim = rvh.HairSim(rvh.default_config(S, N, flags=flags, strands_per_thread=spt))
    for k in env:
        del os.environ[k]
    sim.set_colliders(cols)
    sim.upload(rvh.scenes.synthetic_head(S, N, 2.5))
    sim.step(DT, 0.0)
    sim.step_n(35, DT, 0.1)
    sim.step_n(3, DT, 1.0)
    sim.step(DT, 1.2)
    assert np.isfinite(sim.download()).all()
    if flags & rvh.GRID_ON:
        assert np.abs(sim.download_grid()).max() > 0
    sim.close()
    print("ok late paths", S, N, flags, spt, env, flush=True)
