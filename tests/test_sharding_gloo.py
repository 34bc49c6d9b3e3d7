"""World-size-2 test of the multi-GPU protocol on CPU (gloo): strands sharded by contiguous ranges, every rank
integrates + splats its own strands into a private grid, the int64 grids are all-reduced (sum), every rank gathers
from the reduced grid.  With the oracle standing in for the kernels this must reproduce the single-process step
bit for bit -- the property the NCCL path relies on (integer accumulators => rank-count independent)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, N, L, extra, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import orc
    import rvh_b200 as rvh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = np.float32(1.0 / 60.0)
    cols = rvh.scenes.bench_colliders()
    lo, hi = rvh.scenes.shard_range(S, rank, world)
    st = rvh.scenes.synthetic_head(hi - lo, N, L, first_strand=lo, colliders=cols)     # every rank makes exactly its shard
    rest = np.float32(L) / np.float32(N - 1)
    p = orc.default_params(hi - lo, N, orc.GRID_ON | orc.WIND_B | extra, rest_length=rest)
    for k in range(2):
        T = np.float32(0.25) + np.float32(k) * dt
        st = orc.phase_integrate(p, cols, dt, T, st)
        st, grid = orc.phase_splat(p, dt, st)
        g = torch.from_numpy(grid)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)                                       # the step's one collective
        st = orc.phase_gather(p, st, g.numpy())
    parts = [None] * world
    dist.gather_object((lo, hi, st, g.numpy()), parts if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.concatenate([q[2] for q in sorted(parts, key=lambda q: q[0])])
        np.savez(out_path, state=full, grid=parts[0][3])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("S,N,L,extra", [(3001, 16, 0.4, 0), (2048, 32, 2.5, 0), (2500, 12, 0.4, 128)])     # 128 = ORC_REPULSION_ON: reads the reduced grid too
def test_two_rank_sharded_step_equals_single_process(tmp_path, S, N, L, extra):
    import torch.multiprocessing as tmp
    import orc
    import rvh_b200 as rvh
    out = str(tmp_path / "sharded.npz")
    tmp.spawn(_worker, args=(2, _free_port(), S, N, L, extra, out), nprocs=2, join=True)
    z = np.load(out)
    dt = np.float32(1.0 / 60.0)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, colliders=cols)
    rest = np.float32(L) / np.float32(N - 1)
    p = orc.default_params(S, N, orc.GRID_ON | orc.WIND_B | extra, rest_length=rest)
    for k in range(2):
        st, grid = orc.step(p, cols, dt, np.float32(0.25) + np.float32(k) * dt, st)
    assert np.array_equal(z["grid"], grid), "all-reduced grid differs from the single-process grid"
    assert np.array_equal(z["state"].view(np.uint32), st.view(np.uint32)), "sharded state differs from the single-process state"


def test_shards_cover_and_generate_identically():
    import rvh_b200 as rvh
    S, N = 1000, 8
    full = rvh.scenes.synthetic_head(S, N, 2.5)
    for R in (2, 3, 8):
        parts = []
        for r in range(R):
            lo, hi = rvh.scenes.shard_range(S, r, R)
            parts.append(rvh.scenes.synthetic_head(hi - lo, N, 2.5, first_strand=lo))
        assert np.array_equal(np.concatenate(parts), full)
