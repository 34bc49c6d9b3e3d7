"""Two-GPU test of the sharded path through the C ABI (rvh_create_sharded + NCCL grid all-reduce): the result must
be bit-identical to the one-GPU run of the same strands (integer grid => GPU-count independent)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, uid, S, N, L, steps, out_dir, exchange):
    os.environ["RVH_GRID_EXCHANGE"] = exchange
    sys.path.insert(0, ROOT)
    import rvh_b200 as rvh
    dt = float(np.float32(1.0 / 60.0))
    cols = rvh.scenes.bench_colliders()
    lo, hi = rvh.scenes.shard_range(S, rank, world)
    st = rvh.scenes.synthetic_head(hi - lo, N, L, first_strand=lo, colliders=cols)
    rest = float(np.float32(L) / np.float32(N - 1))
    cfg = rvh.default_config(hi - lo, N, flags=rvh.GRID_ON | rvh.WIND_B | rvh.KEEP_ORDER, device=rank, rest_length=rest)
    sim = rvh.HairSim(cfg, rank=rank, nranks=world, nccl_id=uid)
    mode = sim.exchange_mode()
    sim.set_colliders(cols)
    sim.upload(st)
    for k in range(steps):
        sim.step(dt, 0.1 * k)
    out = sim.download()
    grid = sim.download_grid()
    sim.close()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, state=out, grid=grid, mode=mode)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("S,N,L", [(20000, 16, 0.4), (8192, 32, 2.5), (600000, 8, 0.4)])      # the last one is wide enough for the fused grid clear and the collider mask
def test_two_gpu_sharded_equals_one_gpu(tmp_path, S, N, L, exchange):
    import torch.multiprocessing as tmp
    import rvh_b200 as rvh
    steps = 3
    uid = rvh.nccl_unique_id()
    tmp.spawn(_worker, args=(2, uid, S, N, L, steps, str(tmp_path), exchange), nprocs=2, join=True)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(2)]
    print("grid exchange mode:", [str(p["mode"]) for p in parts])
    if exchange == "nccl":
        assert all(str(p["mode"]) == "nccl-allreduce" for p in parts)
    sharded = np.concatenate([p["state"] for p in parts])
    dt = float(np.float32(1.0 / 60.0))
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    rest = float(np.float32(L) / np.float32(N - 1))
    cfg = rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B | rvh.KEEP_ORDER, rest_length=rest)
    sim = rvh.HairSim(cfg)
    sim.set_colliders(cols)
    sim.upload(st)
    for k in range(steps):
        sim.step(dt, 0.1 * k)
    single = sim.download()
    grid = sim.download_grid()
    sim.close()
    assert np.array_equal(parts[0]["grid"], parts[1]["grid"]), "ranks disagree on the reduced grid"
    assert np.array_equal(parts[0]["grid"], grid), "2-GPU grid differs from the 1-GPU grid"
    assert np.array_equal(sharded.view(np.uint32), single.view(np.uint32)), "2-GPU state differs from the 1-GPU state"


def _lockstep_worker(rank, world, uid, out_dir):
    """Rank 1 never steps: rank 0's fused grid exchange must give up after RVH_EXCHANGE_TIMEOUT_MS and report it."""
    os.environ["RVH_GRID_EXCHANGE"] = "p2p"
    os.environ["RVH_EXCHANGE_TIMEOUT_MS"] = "300"
    sys.path.insert(0, ROOT)
    import time
    import rvh_b200 as rvh
    S, N = 4096, 8
    cols = rvh.scenes.bench_colliders()
    lo, hi = rvh.scenes.shard_range(S, rank, world)
    st = rvh.scenes.synthetic_head(hi - lo, N, 2.5, first_strand=lo, colliders=cols)
    sim = rvh.HairSim(rvh.default_config(hi - lo, N, flags=rvh.GRID_ON, device=rank), rank=rank, nranks=world, nccl_id=uid)
    sim.set_colliders(cols)
    sim.upload(st)
    res = {"mode": sim.exchange_mode(), "error": "", "seconds": 0.0}
    if rank == 0:
        t0 = time.perf_counter()
        try:
            sim.step(float(np.float32(1.0 / 60.0)), 0.0)
            sim.sync()
        except rvh.RvhError as e:
            res["error"] = str(e)
        res["seconds"] = time.perf_counter() - t0
    else:
        time.sleep(3.0)                                        # alive, but not stepping
    np.savez(os.path.join(out_dir, "lock%d.npz" % rank), **res)
    os._exit(0)                                                # the contexts are out of lockstep for good: no orderly teardown


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_a_rank_that_does_not_step_is_an_error_not_a_hang(tmp_path):
    import torch.multiprocessing as tmp
    import rvh_b200 as rvh
    uid = rvh.nccl_unique_id()
    tmp.spawn(_lockstep_worker, args=(2, uid, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(str(tmp_path / "lock0.npz"))
    if str(r0["mode"]) != "peer-memory-fused":
        pytest.skip("peer mapping not available on this box: NCCL path (its own watchdog applies)")
    assert "timed out waiting for rank 1" in str(r0["error"]), str(r0["error"])
    assert float(r0["seconds"]) < 2.5


def _host_worker(rank, world, uid, S, N, L, out_dir):
    os.environ["RVH_GRID_EXCHANGE"] = "p2p"
    sys.path.insert(0, ROOT)
    import rvh_b200 as rvh
    dt = float(np.float32(1.0 / 60.0))
    cols = rvh.scenes.bench_colliders()
    lo, hi = rvh.scenes.shard_range(S, rank, world)
    buf = rvh.scenes.synthetic_head(hi - lo, N, L, first_strand=lo, colliders=cols)
    rest = float(np.float32(L) / np.float32(N - 1))
    sim = rvh.HairSim(rvh.default_config(hi - lo, N, flags=rvh.GRID_ON | rvh.WIND_B, device=rank, rest_length=rest), rank=rank, nranks=world, nccl_id=uid)
    sim.set_colliders(cols)
    sim.step_host(buf, dt, 0.3)                                 # pipelined (>= 128K strands per rank), grid exchange inside
    sim.step_host(buf, dt, 0.3 + dt)
    sim.close()
    np.savez(os.path.join(out_dir, "host%d.npz" % rank), state=buf)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_pipelined_step_host_equals_one_gpu(tmp_path):
    """rvh_step_host's chunk pipeline on sharded contexts: two ranks x 140K strands, two host round trips, against one GPU."""
    import torch.multiprocessing as tmp
    import rvh_b200 as rvh
    S, N, L = 280000, 8, 0.4
    uid = rvh.nccl_unique_id()
    tmp.spawn(_host_worker, args=(2, uid, S, N, L, str(tmp_path)), nprocs=2, join=True)
    sharded = np.concatenate([np.load(str(tmp_path / ("host%d.npz" % r)))["state"] for r in range(2)])
    dt = float(np.float32(1.0 / 60.0))
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    sim = rvh.HairSim(rvh.default_config(S, N, flags=rvh.GRID_ON | rvh.WIND_B, rest_length=float(np.float32(L) / np.float32(N - 1))))
    sim.set_colliders(cols)
    for k in range(2):
        sim.upload(st)
        sim.step(dt, 0.3 + k * dt)
        st = sim.download()
    sim.close()
    assert np.array_equal(sharded[:, 0:2].view(np.uint32), st[:, 0:2].view(np.uint32))
