"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
workload table covers BASELINE.json's configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line_for_the_shipped_scene():
    d = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "strand-point updates/sec" and d["unit"] == "strand-point updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "c1" and d["config"]["points_per_strand"] == 10
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_table_names_every_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    for w in ("c1", "c2", "c3", "c4", "c5", "ns_full", "ns_nogrid", "c3_sdf"):
        S, N, L, flags, desc = bench.WORKLOADS[w]
        assert S >= 900 and N >= 10 and desc
    assert bench.WORKLOADS["c1"][:2] == (900, 10) and bench.WORKLOADS["c5"][:2] == (4000000, 32)
    b1, b2 = bench.bytes_per_strand(32, True)
    assert b1 == 48 * 31 + 12 and b2 == 36 * 31                     # B1 and B2 - B1


def test_reference_arm_processes_share_one_grid_bit_exactly():
    """The reference arm splits one head over one process per core and sums the int32 grids at the shader's splat/gather
    barrier.  Four processes x 512 strands must end, bit for bit, where ONE dispatch of the 2048 strands ends."""
    import numpy as np
    import pytest
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import orc
    import rvh_b200 as rvh
    tag = "N10"
    if not orc.ref_available(tag):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    per, cores, N, L, steps = 512, 4, 10, 2.5, 3
    _, got = bench.ref_multiprocess_run(tag, per, N, L, cores, steps, 0, want_state=True)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(per * cores, N, L, colliders=cols)
    for k in range(steps):
        st, _, _ = orc.ref_dispatch(tag, st, cols, bench.DT, bench.DT * k)
    assert np.array_equal(got.view(np.uint32), st.view(np.uint32))
    assert np.abs(got[:, 1]).max() > 0


def test_spatial_shards_partition_the_head_and_keep_every_strand():
    """bench.py --scaling strong: the ids in Morton order of their roots, cut into `world` pieces.  The pieces are a partition of
    the ids, each is spatially compact (a fraction of the head's bounding box), and a strand generated through `ids=` is the strand
    the contiguous generator makes for that id."""
    import numpy as np
    import rvh_b200 as rvh
    S, R = 20000, 8
    pieces = [rvh.scenes.spatial_shard_ids(S, r, R) for r in range(R)]
    allids = np.sort(np.concatenate(pieces))
    assert np.array_equal(allids, np.arange(S, dtype=np.uint64))
    assert max(len(p) for p in pieces) - min(len(p) for p in pieces) <= 1
    whole = rvh.scenes.synthetic_head(S, 4, 2.5)
    box = np.prod(whole[:, 0, 0, :3].max(0) - whole[:, 0, 0, :3].min(0))
    fracs = []
    for p in pieces:
        part = rvh.scenes.synthetic_head(len(p), 4, 2.5, ids=p)
        assert np.array_equal(part, whole[p.astype(np.int64)])
        r = part[:, 0, 0, :3]
        fracs.append(float(np.prod(r.max(0) - r.min(0)) / box))
    # a contiguous id range covers the whole box (fraction ~1); a Morton range can straddle an octant boundary, so not every piece is small
    assert np.median(fracs) < 0.4 and max(fracs) < 0.8, fracs
