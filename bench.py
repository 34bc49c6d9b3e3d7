#!/usr/bin/env python3
"""bench.py -- strand-point updates/s of the guide-strand physics step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

N>1 is launched by torchrun (one rank per GPU); strands are sharded across ranks (weak scaling:
every rank owns `S` strands) and the voxel grid is all-reduced with NCCL once per step.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what each field means.

--impl reference times the reference's own compute.comp text (oracle/_ref: the shader compiled as C++
against the reference's vendored glm, see oracle/ref_build/) on the host cores, one process per core; if
oracle/_ref is not there it falls back to the OpenMP C port (oracle/oracle.c) and says so.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = float(np.float32(1.0 / 60.0))
UNIT = "strand-point updates/s"

# name -> (strands per GPU, points, strand length, flags, description)
WORKLOADS = {
    # north-star target shape with everything the full step does
    "ns_full": (1 << 20, 32, 2.5, "grid+windB", "1M strands x 32 points per GPU: gravity + wind B + sphere & 5 ellipsoid colliders + voxel-grid friction (int64 grid 64^3)"),
    # north-star "integrate+FTL+collision" target (no hair-hair)
    "ns_nogrid": (1 << 20, 32, 2.5, "windB", "1M strands x 32 points per GPU: gravity + wind B + sphere & 5 ellipsoid colliders, no hair-hair grid"),
    # BASELINE.json configs[0]: the reference's own scene (main.cpp:226-237), 900 guides x 10 points from mannequin_segment.obj
    # follicles (srand(8)), frozen in tests/golden/c1_reference_scene.npz by the reference's own Hair::Hair
    "c1": (900, 10, 2.5, "grid", "C1: the shipped scene, 900 strands x 10 points (Hair::Hair follicles, reference colliders, grid friction as the shader always runs it); 432 KB of state: L2-resident, launch-bound"),
    "c2": (16384, 32, 2.5, "windB", "C2: 16K strands x 32 points, gravity + wind + colliders, no hair-hair (L2-resident, launch-bound)"),
    "c3": (100000, 64, 2.5, "grid", "C3: 100K strands x 64 points, colliders + voxel-grid friction"),
    "c4": (1000000, 16, 0.4, "grid", "C4: 1M fur strands x 16 points per GPU, voxel-grid friction"),
    "c5": (4000000, 32, 2.5, "grid+windB", "C5: 4M strands x 32 points per GPU stress (wind + colliders + grid)"),
    # north-star extensions that the reference does not contain (own oracle modes): head SDF (TMA-staged) + repulsion
    "c3_sdf": (100000, 64, 2.5, "grid+sdf+rep", "C3 as BASELINE.json words it: 100K strands x 64 points, head-SDF collision + voxel-grid friction + repulsion"),
    "ns_sdf": (1 << 20, 32, 2.5, "grid+windB+sdf", "1M strands x 32 points per GPU: gravity + wind B + sphere + head SDF + voxel-grid friction"),
}
SDF_ORIGIN, SDF_EXTENT, SDF_CELL = (-2.0, -2.2, -1.8), (4.0, 6.2, 3.4), 0.05     # lattice over head, neck, bust and shoulders (main.cpp:229-237)


def sdf_lattice():
    import math
    return [int(math.ceil(e / SDF_CELL)) + 1 for e in SDF_EXTENT], np.array(SDF_ORIGIN, np.float32), float(np.float32(SDF_CELL))


def parse_flags(rvh, s):
    f = 0
    if "grid" in s:
        f |= rvh.GRID_ON
    if "windB" in s:
        f |= rvh.WIND_B
    if "windA" in s:
        f |= rvh.WIND_A
    if "sdf" in s:
        f |= rvh.SDF_ON
    if "sdftma" in s:
        f |= rvh.SDF_TMA
    if "rep" in s:
        f |= rvh.REPULSION_ON
    return f


def bytes_per_strand(N, grid):
    """Algorithmic bytes per strand per step (fp32 xyz only; SURVEY.md 8d / DESIGN.md): K1 reads p,v of N-1 points +
    root p and writes p,v of N-1 points; the reference-shaped full step re-reads p,v and re-writes v for the gather."""
    b1 = 48 * (N - 1) + 12
    b2 = 36 * (N - 1)
    return b1, (b2 if grid else 0)


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled from a thread (`nvidia-smi -lms 200`, the recipe's
    line, is too coarse for a timed region of a quarter of a second); falls back to one nvidia-smi query.  The period
    matters: NVML queries are not free for the GPU -- polling every 2 ms slowed the step by 3.6 % (splat 0.425 -> 0.467 ms),
    so the poll runs every 25 ms and the cost that remains is reported (`roofline.sampler_overhead_ms_per_step`)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.period = float(os.environ.get("RVH_BENCH_SAMPLER_PERIOD", "0.025"))

    def _sample(self):
        import pynvml as nv
        h = self.h
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        for bit, nm in names.items():
            if r & bit:
                self.reasons.add(nm)
        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)

    def _poll(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                break
            self.stop_flag.wait(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[0].isdigit() else self.gpu
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self):
        if self.thread is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [], "how": "nvidia-smi after the run (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        try:
            self._sample()                      # one more right at the end of the timed region (short runs: K steps may last < one period)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "how": "NVML polled every %d ms during the timed region" % int(self.period * 1e3)}


# ---- the reference on the host cores ----------------------------------------------------------------------

def c1_scene():
    """(state0 [900,3,10,4], colliders [6,48]) of the shipped scene, as the reference's own constructors produced them."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_reference_scene.npz"))
    return np.ascontiguousarray(g["state0"], np.float32), np.ascontiguousarray(g["colliders"], np.float32)


def _ref_tag(N, flags_s):
    import orc
    tag = "N%d%s" % (N, "_windB" if "windB" in flags_s else ("_windA" if "windA" in flags_s else ""))
    return tag if orc.ref_available(tag) else None


def vulkan_probe():
    """BASELINE.json's intended CPU baseline is the unmodified compute.comp on Mesa lavapipe.  Probed at every start of the
    reference arm: an ICD manifest, a Vulkan loader and a GLSL compiler are all needed to run it."""
    import ctypes.util
    import glob
    import shutil
    icds = [f for d in ("/usr/share/vulkan/icd.d", "/etc/vulkan/icd.d", os.path.expanduser("~/.local/share/vulkan/icd.d")) for f in glob.glob(os.path.join(d, "*.json"))]
    icds += [f for f in os.environ.get("VK_ICD_FILENAMES", "").split(":") if f and os.path.exists(f)]
    loader = ctypes.util.find_library("vulkan")
    compiler = shutil.which("glslangValidator") or shutil.which("glslc")
    lavapipe = [f for f in icds if "lvp" in os.path.basename(f) or "lavapipe" in os.path.basename(f)]
    ok = bool(lavapipe and loader and compiler)
    return {"available": ok, "icd_manifests": icds, "loader": loader, "glsl_compiler": compiler,
            "note": "lavapipe + loader + GLSL compiler present: the shader could run through Vulkan, but no harness for it ships here" if ok else
                    "compute.comp on Mesa lavapipe (BASELINE.json's intended CPU baseline) cannot run in this image: %s; the reference's shader text is compiled "
                    "as C++ against its vendored glm instead (oracle/_ref)" % ", ".join(
                        m for m, missing in (("no lavapipe ICD manifest", not lavapipe), ("no libvulkan", not loader), ("no glslangValidator/glslc", not compiler)) if missing)}


_HOOK_T = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_size_t)


def _ref_worker(tag, idx, cores, S, N, L, first, steps, warmup, barrier, partial_raw, total_raw, cells4, q, want_state=False):
    """One process = one host core = one dispatch of the reference shader text over its own strand range of ONE head: at the
    shader's third barrier (compute.comp:255, splat done / gather next) the processes sum their int32 grids through shared
    memory, exactly what one dispatch over all strands accumulates (int32, wrapping), so the gather sees every strand."""
    import orc
    import rvh_b200 as rvh
    lib = orc.ref_compute(tag)
    partial = np.frombuffer(partial_raw, np.int32).reshape(cores, cells4)
    total = np.frombuffer(total_raw, np.int32)
    lo, hi = (cells4 * idx) // cores, (cells4 * (idx + 1)) // cores

    def hook(k, grid_ptr, nbytes):
        if k != 3:
            return
        g = np.ctypeslib.as_array(C.cast(grid_ptr, C.POINTER(C.c_int32)), shape=(cells4,))
        partial[idx, :] = g
        barrier.wait()
        total[lo:hi] = partial[:, lo:hi].sum(axis=0, dtype=np.int32)          # int32 wrap-around = the shader's atomicAdd
        barrier.wait()
        g[:] = total

    cb = _HOOK_T(hook)
    if cores > 1:
        lib.ref_set_barrier_hook(cb)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L, first_strand=first, colliders=cols)
    for w in range(warmup):
        st, _, _ = orc.ref_dispatch(tag, st, cols, DT, DT * w)
    barrier.wait()
    t0 = time.perf_counter()
    for k in range(steps):
        st, _, _ = orc.ref_dispatch(tag, st, cols, DT, DT * (warmup + k))
    el = time.perf_counter() - t0
    q.put((el, idx, st if want_state else None))


def ref_multiprocess_run(tag, per, N, L, cores, steps, warmup, want_state=False):
    """`cores` processes x `per` strands of one synthetic head through the reference shader text, grids shared (see _ref_worker).
    Returns (seconds of the timed steps = max over the processes, final Strand[cores*per] or None)."""
    import orc
    G = orc.ref_compute(tag).ref_shader_grid_dim()
    cells4 = G * G * G * 4
    ctx = mp.get_context("fork")
    partial_raw = ctx.RawArray("i", cores * cells4)
    total_raw = ctx.RawArray("i", cells4)
    q, bar = ctx.Queue(), ctx.Barrier(cores)
    procs = []
    for i in range(cores):
        pr = ctx.Process(target=_ref_worker, args=(tag, i, cores, per, N, L, i * per, steps, warmup, bar, partial_raw, total_raw, cells4, q, want_state))
        pr.start()
        procs.append(pr)
    res = sorted([q.get() for _ in procs], key=lambda r: r[1])
    for pr in procs:
        pr.join()
    state = np.concatenate([r[2] for r in res]) if want_state else None
    return max(r[0] for r in res), state


REF_MAX_STRANDS = 1 << 20      # the reference arm's bounded sample: at most this many strands per step (the whole workload at N = 1 GPU)


def time_reference(workload, flags_s, steps, warmup, max_strands=REF_MAX_STRANDS, S_total=None):
    """Times the reference's CPU implementation of the path.  Returns (value, seconds, info dict)."""
    import orc
    import rvh_b200 as rvh
    S_full, N, L, _, _ = WORKLOADS[workload]
    S_full = S_total or S_full
    cores = os.cpu_count() or 1
    tag = _ref_tag(N, flags_s) if abs(L - 2.5) < 1e-6 else None      # the shader hard-codes strand length 2.5 (compute.comp:139)
    if "sdf" in flags_s or "rep" in flags_s:
        tag = None                                                   # extensions do not exist in the shader: only the C port has them
    if workload == "c1" and tag is not None:
        # the whole workload is 900 strands: one dispatch of the reference shader text per step, one core, exactly as shipped
        orc.ref_compute(tag)
        st, cols = c1_scene()
        for w in range(warmup):
            st, _, _ = orc.ref_dispatch(tag, st, cols, DT, DT * w)
        t0 = time.perf_counter()
        for k in range(steps):
            st, _, _ = orc.ref_dispatch(tag, st, cols, DT, DT * (warmup + k))
        el = time.perf_counter() - t0
        info = {"kind": "reference", "cores": 1, "strands_measured": 900,
                "sample": "the full workload: 900 strands x 10 points per step, %d steps: the reference's compute.comp text compiled as C++ against its vendored glm "
                          "(oracle/_ref/libref_compute_%s.so), one dispatch per step on one core, -O2" % (steps, tag)}
        return 900 * N * steps / el, el, info
    if tag is not None:
        # the shader TU keeps its buffers in globals, so each core runs its own process over its own strand range of the same
        # head; the int32 grids are summed across the processes before the gather (see _ref_worker): no work is skipped
        orc.ref_compute(tag)                                         # also maps oracle/_ref into THIS process (the forked workers inherit it)
        S = min(S_full, max_strands)
        per = (S + cores - 1) // cores
        S = per * cores if per * cores <= S_full else S
        per = S // cores
        S = per * cores
        el, _ = ref_multiprocess_run(tag, per, N, L, cores, steps, warmup)
        info = {"kind": "reference", "cores": cores, "strands_measured": S,
                "sample": "%d strands x %d points per step%s, %d steps after %d warm-up: the reference's compute.comp text compiled as C++ against its vendored glm "
                          "(oracle/_ref/libref_compute_%s.so, -O2; no Vulkan/lavapipe in this image), one process per host core over its own strand range of one head, "
                          "the int32 grids summed across the processes at the shader's splat/gather barrier so that every strand meets every other in the grid"
                          % (S, N, " (the whole workload)" if S == S_full else " (bounded sample of the %d-strand workload; throughput is linear in strands)" % S_full, steps, warmup, tag)}
        return S * N * steps / el, el, info
    # fall back to the OpenMP C port
    S = min(S_full, max_strands)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L)
    rest = np.float32(L) / np.float32(N - 1)
    of = (orc.GRID_ON if "grid" in flags_s else 0) | (orc.WIND_B if "windB" in flags_s else 0) | (orc.WIND_A if "windA" in flags_s else 0)
    of |= (orc.SDF_ON if "sdf" in flags_s else 0) | (orc.REPULSION_ON if "rep" in flags_s else 0)
    if "sdf" in flags_s:
        dim, origin, cell = sdf_lattice()
        orc.set_head_sdf(orc.sdf_bake_colliders(cols, dim, origin, cell), origin, cell)
    p = orc.default_params(S, N, of, rest_length=rest)
    grid = orc.new_grid(p)
    L_ = orc.lib()
    threads = L_.orc_max_threads()
    colp = np.ascontiguousarray(cols, np.float32)

    def one(t):
        L_.orc_step_parallel(C.byref(p), orc._f(colp), DT, t, orc._f(st), grid.ctypes.data_as(orc._i64p), threads)

    for w in range(warmup):
        one(DT * w)
    t0 = time.perf_counter()
    for k in range(steps):
        one(DT * (warmup + k))
    el = time.perf_counter() - t0
    info = {"kind": "port", "cores": threads, "strands_measured": S,
            "sample": "%d strands x %d points per step, %d steps: C restatement of compute.comp (oracle/oracle.c, OpenMP over strands); "
                      "oracle/_ref has no build of this variant (strand length %.2f / N=%d%s)" % (S, N, steps, L, N, "; extension flags" if ("sdf" in flags_s or "rep" in flags_s) else "")}
    return S * N * steps / el, el, info


def workload_config(args, world):
    """The `config` object of the JSON line: the workload only, identical for both arms (--impl b200 / reference)."""
    S, N, L, flags_s, desc = WORKLOADS[args.workload]
    if args.flags is not None:
        flags_s = args.flags
        desc += " [flags overridden: %s]" % flags_s
    S_total = S * world if args.scaling == "weak" else S
    state_mb = (S_total // world if args.scaling == "strong" else S) * N * 24 / 1e6
    return {"workload": args.workload, "description": desc, "features": flags_s, "strands_per_gpu": S if args.scaling == "weak" else S_total // world,
            "strands_total": S_total, "points_per_strand": N, "strand_length": L, "dt": DT, "scaling": args.scaling,
            "l2": "state %.0f MB per GPU > 126 MB L2, no flush needed" % state_mb if state_mb > 126 else "state %.1f MB is L2-resident (launch/latency-bound config)" % state_mb}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0                                           # rank 0 alone runs the CPU arm
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = workload_config(args, world)
    steps = max(1, min(args.steps, 40))                    # 1M x 32 takes ~1.4 s per step on 16 cores: the default --steps 300 is capped, the line says what ran
    warmup = max(1, min(args.warmup, 10))
    val, el, info = time_reference(args.workload, cfg["features"], steps, warmup, S_total=cfg["strands_total"])
    out = {"impl": "reference", "metric": "strand-point updates/sec", "value": val, "unit": UNIT,
           "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * el / steps,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": cfg,
           "cpu_baseline": dict(info, value=val, unit=UNIT, lavapipe=vulkan_probe()),
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


class Ranks:
    """torch.distributed plumbing (NCCL) for the N > 1 launch; every rank drives its own context."""

    def __init__(self, rvh):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.numa = self._bind_to_gpu_numa_node() if self.world > 1 else "1 GPU: not bound"
        self.dist = None
        self.rvh = rvh
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def _bind_to_gpu_numa_node(self):
        """Multi-GPU e2e is host-bound: every rank streams 2 x 50 GB/s through pinned host memory.  Run each rank (and so allocate its
        pinned buffers) on the NUMA node its GPU hangs off, so that the copies do not cross the socket interconnect."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.local]) if vis and vis.split(",")[0].isdigit() else self.local
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bdf = bus.lower()[-12:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
            if node < 0:
                return "GPU %s reports no NUMA node: not bound" % bdf
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if not cpus:
                return "NUMA node %d of GPU %s has no CPU this process may use: not bound" % (node, bdf)
            os.sched_setaffinity(0, cpus)
            return "rank %d bound to NUMA node %d (%d CPUs) of GPU %s" % (self.rank, node, len(cpus), bdf)
        except Exception as e:                                        # no NVML / no sysfs: run unbound
            return "not bound (%s)" % type(e).__name__

    def new_nccl_id(self):
        """A fresh ncclUniqueId from rank 0 (one per sharded context)."""
        if self.dist is None:
            return None
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            idt = self.torch.tensor(list(self.rvh.nccl_unique_id()), dtype=self.torch.uint8, device="cuda")
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def barrier(self, sim=None):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        if sim is not None:
            sim.sync()

    def max(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_u64(self, x):
        """Every rank's 64-bit word, on every rank."""
        if self.dist is None:
            return [int(x)]
        t = self.torch.tensor([np.int64(np.uint64(x).astype(np.int64))], dtype=self.torch.int64, device="cuda")
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [int(np.uint64(np.int64(o.item()))) for o in out]


def verification_checksum(R, rvh):
    """A fixed scene -- 262,144 strands x 16 points, grid + wind B, 3 steps -- SHARDED over however many ranks run, reduced to
    two 64-bit words: the wrapping sum of the reduced int64 grid and the XOR of every position/velocity bit.  The grid exchange is
    integer and the strands are otherwise independent, so the words are the same for 1, 2, 4 and 8 GPUs: anyone can compare the
    SCALE lines."""
    S_total, N, L = 262144, 16, 2.5
    lo, hi = rvh.scenes.shard_range(S_total, R.rank, R.world)
    cols = rvh.scenes.bench_colliders()
    cfg = rvh.default_config(hi - lo, N, flags=rvh.GRID_ON | rvh.WIND_B, device=R.local, rest_length=float(np.float32(L) / np.float32(N - 1)))
    sim = rvh.HairSim(cfg, rank=R.rank, nranks=R.world, nccl_id=R.new_nccl_id())
    sim.set_colliders(cols)
    sim.init_synthetic_head(lo, L, 8)
    for k in range(3):
        sim.step(DT, 0.1 + k * DT)
    grid = sim.download_grid()                                   # collective: all-reduces on demand
    st = sim.download()
    mode = sim.exchange_mode()
    sim.close()
    words = np.ascontiguousarray(st[:, 0:2]).view(np.uint64)
    x = np.bitwise_xor.reduce(words.reshape(-1)) if words.size else np.uint64(0)
    xs = R.gather_u64(x)
    state_xor = 0
    for v in xs:
        state_xor ^= v
    grid_sum = int(grid.astype(np.int64).view(np.uint64).sum(dtype=np.uint64))
    return {"grid_sum_u64": "%016x" % grid_sum, "state_xor_u64": "%016x" % state_xor, "exchange": mode,
            "what": "262144 strands x 16 points (grid + wind B), sharded over the %d rank(s), 3 steps: wrapping sum of the reduced int64 grid, XOR of all position/velocity "
                    "words; identical for any GPU count" % R.world}


def measure(R, rvh, args, workload, steps, warmup, scaling, full):
    """One workload on the ranks of R: device-resident timed region (+ per-kernel split, + the end-to-end legs when `full`)."""
    torch = R.torch
    S, N, L, flags_s, desc = WORKLOADS[workload]
    if full and args.flags is not None:
        flags_s = args.flags
    S_total = S * R.world if scaling == "weak" else S
    ids = None
    if scaling == "strong":
        lo, hi = rvh.scenes.shard_range(S_total, R.rank, R.world)
        first_strand, S = lo, hi - lo
    else:
        first_strand = R.rank * S
    device_init = args.device_init or S * N >= (1 << 26) or not full
    if scaling == "strong" and R.world > 1 and workload != "c1" and os.environ.get("RVH_BENCH_STRONG_SHARDS", "spatial") == "spatial":
        # ONE head over several GPUs: spatial domain decomposition (the ids sorted by the Morton code of their roots, cut into world
        # pieces).  Contiguous id ranges would give every rank a uniformly thinned copy of the whole head (the generator hashes the id).
        ids = rvh.scenes.spatial_shard_ids(S_total, R.rank, R.world, colliders=rvh.scenes.bench_colliders())
        device_init = False
    flags = parse_flags(rvh, flags_s)
    grid_on = bool(flags & rvh.GRID_ON)
    rest = float(np.float32(L) / np.float32(N - 1))
    cols = rvh.scenes.bench_colliders()
    c1 = workload == "c1"
    if c1:
        if R.world > 1 and scaling != "strong":
            raise SystemExit("c1 is one fixed scene of 900 strands: use --scaling strong to shard it")
        c1_state, cols = c1_scene()
        device_init = False

    aos_bytes = S * 48 * N
    pinned = host = None
    if full or c1 or ids is not None:
        # synthetic inputs in pinned host memory (global strand ids => every rank makes its own shard)
        pinned = torch.empty(aos_bytes // 4, dtype=torch.float32, pin_memory=True)
        host = pinned.numpy().reshape(S, 3, N, 4)
        if c1:
            host[:] = c1_state[first_strand:first_strand + S]
        elif not device_init:
            rvh.scenes.synthetic_head(S, N, L, first_strand=first_strand, colliders=cols, out=host, ids=ids)

    cfg = rvh.default_config(S, N, flags=flags, device=R.local, rest_length=rest, strands_per_thread=args.spt)
    sim = rvh.HairSim(cfg, rank=R.rank, nranks=R.world, nccl_id=R.new_nccl_id())
    sim.set_colliders(cols)
    if flags & rvh.SDF_ON:
        dim, origin, cell = sdf_lattice()
        sim.bake_head_sdf_from_colliders(dim, origin, cell)     # GPU bake of the scene's own ellipsoids
    if device_init:
        sim.init_synthetic_head(first_strand, L, 8)
        if pinned is not None:
            sim.download_ptr(pinned.data_ptr(), aos_bytes)              # the e2e leg starts from the same state in host memory
    else:
        sim.upload_ptr(pinned.data_ptr(), aos_bytes)
    sim.sync()
    exchange = sim.exchange_mode()
    sdf_mode = sim.sdf_mode()

    # small scenes: rvh_step_n's fast paths (many steps per launch / CUDA-graph replay) are what a user gets; they carry no
    # per-kernel events, so the roofline kernel's duration is the step itself there
    S_pad = (S + 255) // 256 * 256
    fast = None
    if R.world == 1 and not (flags & rvh.SDF_ON):
        if not grid_on and S * N * 24 <= 126e6:
            fast = "up to 32 steps per launch (k_ftl_wave: wavefront over the steps on small scenes; k_ftl_step MULTI otherwise)"
        elif grid_on and not (flags & rvh.REPULSION_ON) and 2 * (-(-(S_pad // (args.spt if args.spt in (1, 2) else (2 if S >= 131072 else 1))) // 128)) <= \
                torch.cuda.get_device_properties(R.local).multi_processor_count and os.environ.get("RVH_SCENE_CTAS", "2") != "0":
            fast = "up to 32 whole steps per persistent cooperative launch (k_scene_step: FTL without gather || clear | splat | per-point gather from the int64 accumulators, grid barriers in between)"
        elif grid_on and "wind" not in flags_s and S_pad * N <= (1 << 23):
            fast = "CUDA-graph replay of the step"
    small = fast is not None
    kernel_events = not (args.no_kernel_events or small)

    # ---- device-resident timed region -------------------------------------------------------
    sim.step_n(max(warmup, 3), DT, 0.0, timed=True)
    sim.profile_enable(2 if kernel_events else 0)               # timed region: events around the roofline kernel only
    sim.profile_read()
    launches0 = sim.kernel_launches()
    sampler = ClockSampler(R.local)
    if R.rank == 0 and full:
        sampler.start()
    R.barrier(sim)
    ms = sim.step_n(steps, DT, DT * warmup, timed=True)         # CUDA events on the context's stream
    R.barrier(sim)
    clocks = sampler.stop() if (R.rank == 0 and full) else None
    prof = sim.profile_read()
    sim.profile_enable(0)
    launches = sim.kernel_launches() - launches0
    n2 = min(steps, 50)
    ms_plain = sim.step_n(n2, DT, DT * (warmup + steps), timed=True) / n2      # same steps with neither the clock sampler nor kernel events running
    # the split over ALL kernels comes from a separate, untimed pass right after the timed region (one launch per kernel and step)
    k1_timed = prof["ftl_step"]
    sim.profile_enable(1)
    sim.step_n(n2, DT, DT * (warmup + steps + n2), timed=True)
    prof = sim.profile_read()
    sim.profile_enable(0)
    if kernel_events:
        prof["ftl_step"] = k1_timed                             # the roofline kernel: measured inside the timed region
    ms = R.max(ms)
    value = S_total * N * steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak = 6650.0; peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    b1, b2 = bytes_per_strand(N, grid_on)
    per_kernel = {k: (v["ms"] / v["launches"] if v["launches"] else 0.0) for k, v in prof.items()}
    dominant = max(per_kernel, key=lambda k: per_kernel[k])
    # graph replay: the kernel is the same as in the per-kernel pass, its duration comes from there; many-steps-per-launch: the
    # launch IS the steps, so the step time of the timed region is the per-step duration of the kernel
    scene = bool(fast) and fast.startswith("up to 32 whole steps")          # k_scene_step: the launch IS the steps, all phases inside
    k1_ms = per_kernel["ftl_step"] if (kernel_events or (small and grid_on and not scene and per_kernel["ftl_step"])) else ms / steps
    # every kernel of the step with its ALGORITHMIC bytes per launch (fp32 xyz only) and CUDA-event duration
    kernels = {"k_ftl_step": {"what": "integrate + collide + FTL + corrected velocity%s" % (" + fused gather of the previous grid" if grid_on else ""),
                              "algorithmic_bytes_per_launch": S * b1, "avg_launch_ms": k1_ms, "bound": "hbm"}}
    if scene:
        kernels = {"k_scene_step": {"what": "the whole step in one persistent launch: integrate + collide + FTL || grid clear | splat | gather from the int64 accumulators",
                                    "algorithmic_bytes_per_launch": S * (b1 + b2), "avg_launch_ms": k1_ms, "bound": "hbm (scored against it; the launch is latency-bound: three grid barriers and a 9-row dependent chain)"}}
    elif grid_on and per_kernel.get("grid_splat"):
        # reads p, v of every moving point once; its time is integer work (32 float->int truncations + adds per point), not bytes
        kernels["k_grid_splat"] = {"what": "corrected velocities -> int64 voxel grid (compute.comp:231-252)", "algorithmic_bytes_per_launch": S * (N - 1) * 24,
                                   "avg_launch_ms": per_kernel["grid_splat"], "bound": "issue (scored against hbm: the roofline the contract allows)"}
    for k in kernels.values():
        k["achieved"] = k["algorithmic_bytes_per_launch"] / (k["avg_launch_ms"] * 1e-3) / 1e9
        k["frac"] = k["achieved"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["avg_launch_ms"])          # the dominant kernel: the longest one
    D = kernels[dom]
    roofline = {"bound": "hbm", "kernel": "%s (%s)" % (dom, D["what"]),
                "achieved": D["achieved"], "peak": peak, "unit": "GB/s", "frac": D["frac"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": D["algorithmic_bytes_per_launch"], "avg_launch_ms": D["avg_launch_ms"],
                "kernels": kernels, "north_star_kernel": "k_scene_step" if scene else "k_ftl_step", "north_star_frac": kernels["k_scene_step" if scene else "k_ftl_step"]["frac"],
                "per_kernel_ms": per_kernel,
                "per_kernel_ms_source": ("ftl_step: CUDA events inside the timed region; " if kernel_events else
                                         "small scene: rvh_step_n replays the step as a CUDA graph (avg_launch_ms: k_ftl_step in the per-kernel pass) or runs many steps per launch (k_ftl_wave / k_ftl_step MULTI / k_scene_step; avg_launch_ms: step time of the timed region; the per-kernel pass below then shows the launch-per-kernel path, not the timed one); ") +
                                        "per_kernel_ms: events around every kernel in a separate pass of %d single steps right after it" % n2,
                "longest_kernel": dominant, "sampler_overhead_ms_per_step": max(0.0, ms / steps - ms_plain),
                "step_bytes": S * (b1 + b2), "step_frac": (S * (b1 + b2) / (ms / steps * 1e-3) / 1e9) / peak,
                "note": "the object describes the DOMINANT (longest) kernel of the step: achieved = its algorithmic bytes / its CUDA-event time. With the grid on that is "
                        "k_grid_splat, which is issue-bound integer work (DESIGN.md section 4), so its HBM fraction is low by construction; the north-star target "
                        "(>= 0.60 of HBM on integrate + FTL + collision) is north_star_frac = k_ftl_step: S*(48*(N-1)+12) bytes / its time inside the timed region. "
                        "step_frac = reference-shaped step bytes S*(84*(N-1)+12) / whole step time"}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            tj = json.load(open(traffic_path))
            roofline["traffic"] = tj.get(workload, {}).get(dom)
            roofline["traffic_per_kernel"] = tj.get(workload)
            roofline["traffic_source"] = "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this command (%s)" % tj.get("_source", "profiles/")
        except Exception:
            pass
    out = {"workload": workload, "value": value, "ms_per_step": ms / steps, "steps": steps, "roofline": roofline, "gpu_launches": int(launches),
           "clocks": clocks, "S": S, "S_total": S_total, "N": N, "flags_s": flags_s, "cols_nbytes": int(cols.nbytes),
           "implementation": {"scene_init": "reference scene frozen from Hair::Hair (tests/golden/c1_reference_scene.npz) + upload" if c1 else ("GPU (rvh_init_synthetic_head)" if device_init else "host (scenes.synthetic_head) + upload"),
                              "parallelism": "strand-sharded x%d%s, grid exchange per step: %s" % (R.world, " (spatial shards: ids in Morton order of their roots)" if ids is not None else "", exchange) if R.world > 1 else "1 GPU", "numa": R.numa,
                              "strands_per_thread": int(sim.cfg.strands_per_thread),
                              "step_n_fast_path": fast or "none",
                              "head_sdf": ("%s lattice, cell %.3f, sampled through %s" % ("x".join(str(d) for d in sdf_lattice()[0]), SDF_CELL, sdf_mode)) if flags & rvh.SDF_ON else None}}
    if not full:
        sim.close()
        return out

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------
    if not args.no_e2e:
        R.barrier(sim)
        sim.step_host_ptr(pinned.data_ptr(), aos_bytes, DT, 0.0)            # warm-up
        R.barrier(sim)
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            sim.step_host_ptr(pinned.data_ptr(), aos_bytes, DT, DT * k)      # H2D Strand[S] + step + D2H Strand[S]
        R.barrier(sim)
        el = R.max(time.perf_counter() - t0)
        out["e2e"] = {"value": S_total * N * args.e2e_steps / el, "unit": UNIT,
                      "h2d_bytes_per_step": S * 32 * N + cols.nbytes + 8, "d2h_bytes_per_step": S * 32 * N,
                      "steps": args.e2e_steps, "what": "rvh_step_host: upload curvePoints+curveVels of the host Strand[S] AoS (pinned), one step, download them back, "
                                                       "every step (correctionVecs are dead across steps and stay on the host); above 128K strands the call pipelines "
                                                       "chunked copies in both PCIe directions against the kernels (positions return before the grid is complete)"}
        # the reference's own per-frame contract: state stays on the GPU, only Time + Collider UBOs go in (Scene.cpp:78-87,133)
        R.barrier(sim)
        n_res = min(steps, 50)
        t0 = time.perf_counter()
        for k in range(n_res):
            sim.set_colliders(cols)
            sim.step(DT, DT * k)
            sim.draw_indirect()
        R.barrier(sim)
        el = R.max(time.perf_counter() - t0)
        out["e2e_resident"] = {"value": S_total * N * n_res / el, "unit": UNIT,
                               "h2d_bytes_per_step": cols.nbytes + 8, "d2h_bytes_per_step": 16, "steps": n_res,
                               "what": "per-frame API as the reference drives it: rvh_set_colliders (UBO) + rvh_step + rvh_draw_indirect read-back; strand state resident"}
    if args.expand:
        sim.expand(12, 42, download=False)                                   # allocation + tables
        times = [sim.expand(12, 42, download=False)[2] for _ in range(5)]
        verts = S * 12 * 43
        ems = statistics.median(times)
        out["expand"] = {"isolines": 12, "divisions": 42, "vertices": verts, "ms": ems, "vertices_per_s": verts / (ems * 1e-3),
                         "write_GBps": verts * 32 / (ems * 1e-3) / 1e9, "what": "k_expand_strands: 2 x float4 per vertex written, guide positions read once (per GPU)"}
    sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="ns_full", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spt", type=int, default=0, help="strands per thread (0 auto)")
    ap.add_argument("--flags", default=None, help="override the workload's feature flags, e.g. grid+windB (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE.json configurations appended as `configs`")
    ap.add_argument("--no-checksum", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: every rank owns the workload's strand count; strong: the count is split over the ranks")
    ap.add_argument("--no-kernel-events", action="store_true", help="time the steps without the per-kernel CUDA events (no roofline per-kernel split)")
    ap.add_argument("--expand", action="store_true", help="also time the guide -> render strand expansion (hair.tesc/hair.tese, 12 isolines x 42 divisions) of the final state")
    ap.add_argument("--device-init", action="store_true", help="generate the synthetic head on the GPU (rvh_init_synthetic_head) instead of uploading it")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import rvh_b200 as rvh
    R = Ranks(rvh)
    m = measure(R, rvh, args, args.workload, args.steps, args.warmup, args.scaling, full=True)

    # ---- the other BASELINE.json configurations, time-capped, beside the headline ------------------
    configs = None
    if not args.no_configs and args.workload == "ns_full" and args.flags is None:
        configs = []
        if R.world == 1:
            plan = [("c1", "weak"), ("c2", "weak"), ("c3", "weak"), ("c4", "weak"), ("c5", "weak")]
        else:
            plan = [("c4", "weak"), ("c5", "strong")]          # configs[3] and configs[4]: the multi-GPU ones
        for w, sc in plan:
            k = 100 if w in ("c1", "c2", "c3") else (40 if w == "c4" else 20)
            c = measure(R, rvh, args, w, k, 3, sc, full=False)
            r = c["roofline"]
            configs.append({"workload": w, "scaling": sc, "strands_total": c["S_total"], "points_per_strand": c["N"], "features": c["flags_s"],
                            "value": c["value"], "unit": UNIT, "ms_per_step": c["ms_per_step"], "steps": k, "dominant_kernel": r["kernel"].split(" ")[0], "roofline_frac": r["frac"], "ftl_frac": r["north_star_frac"],
                            "step_frac": r["step_frac"], "per_kernel_ms": r["per_kernel_ms"], "gpu_launches": c["gpu_launches"],
                            "step_n_fast_path": c["implementation"]["step_n_fast_path"], "parallelism": c["implementation"]["parallelism"]})
    checksum = None if args.no_checksum else verification_checksum(R, rvh)

    cpu = None
    if R.rank == 0 and R.world == 1 and not args.no_cpu_baseline:
        val, el, info = time_reference(args.workload, m["flags_s"], steps=3, warmup=1)
        cpu = dict(info, value=val, unit=UNIT, seconds=el)

    if R.dist is not None:
        R.dist.barrier()
        R.dist.destroy_process_group()
    if R.rank != 0:
        return 0
    out = {
        "metric": "strand-point updates/sec", "value": m["value"], "unit": UNIT, "n_gpus": R.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, R.world), "implementation": m["implementation"],
        "roofline": m["roofline"], "cpu_baseline": cpu, "e2e": m.get("e2e"), "e2e_resident": m.get("e2e_resident"),
        "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "checksum": checksum, "configs": configs,
    }
    if "expand" in m:
        out["expand"] = m["expand"]
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
