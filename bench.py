#!/usr/bin/env python3
"""bench.py -- strand-point updates/s of the guide-strand physics step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

N>1 is launched by torchrun (one rank per GPU); strands are sharded across ranks (weak scaling:
every rank owns `S` strands) and the voxel grid is all-reduced with NCCL once per step.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what each field means.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = float(np.float32(1.0 / 60.0))

# name -> (strands per GPU, points, strand length, flags, description)
WORKLOADS = {
    # north-star target shape with everything the full step does
    "ns_full": (1 << 20, 32, 2.5, "grid+windB", "1M strands x 32 points per GPU: gravity + wind B + sphere & 5 ellipsoid colliders + voxel-grid friction (int64 grid 64^3)"),
    # north-star "integrate+FTL+collision" target (no hair-hair)
    "ns_nogrid": (1 << 20, 32, 2.5, "windB", "1M strands x 32 points per GPU: gravity + wind B + sphere & 5 ellipsoid colliders, no hair-hair grid"),
    "c2": (16384, 32, 2.5, "windB", "C2: 16K strands x 32 points, gravity + wind + colliders, no hair-hair (L2-resident, launch-bound)"),
    "c3": (100000, 64, 2.5, "grid", "C3: 100K strands x 64 points, colliders + voxel-grid friction"),
    "c4": (1000000, 16, 0.4, "grid", "C4: 1M fur strands x 16 points per GPU, voxel-grid friction"),
    "c5": (4000000, 32, 2.5, "grid+windB", "C5: 4M strands x 32 points per GPU stress (wind + colliders + grid)"),
}


def parse_flags(rvh, s):
    f = 0
    if "grid" in s:
        f |= rvh.GRID_ON
    if "windB" in s:
        f |= rvh.WIND_B
    if "windA" in s:
        f |= rvh.WIND_A
    return f


def bytes_per_strand(N, grid):
    """Algorithmic bytes per strand per step (fp32 xyz only): K1 reads p,v of N-1 points +
    root p and writes p,v of N-1 points; with the grid K2 re-reads p,v and re-writes v."""
    b1 = 48 * (N - 1) + 12
    b2 = 36 * (N - 1)
    return b1, (b2 if grid else 0)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(workload, flags_s, seconds=12.0, max_steps=4, sample_strands=131072):
    """The CPU oracle (oracle/liboracle.so, OpenMP over all host threads) on a bounded sample of
    the same workload.  Reported beside the GPU number; never on the product path."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    import rvh_b200 as rvh
    S_full, N, L, _, _ = WORKLOADS[workload]
    S = min(S_full, sample_strands)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L)
    rest = np.float32(L) / np.float32(N - 1)
    of = (orc.GRID_ON if "grid" in flags_s else 0) | (orc.WIND_B if "windB" in flags_s else 0) | (orc.WIND_A if "windA" in flags_s else 0)
    p = orc.default_params(S, N, of, rest_length=rest)
    grid = orc.new_grid(p)
    threads = orc.lib().orc_max_threads()
    L_ = orc.lib()
    colp = np.ascontiguousarray(cols, np.float32)
    L_.orc_step_parallel(C.byref(p), orc._f(colp), DT, 0.0, orc._f(st), grid.ctypes.data_as(orc._i64p), threads)  # warm-up
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and (time.perf_counter() - t0) < seconds:
        L_.orc_step_parallel(C.byref(p), orc._f(colp), DT, DT * (steps + 1), orc._f(st), grid.ctypes.data_as(orc._i64p), threads)
        steps += 1
    el = time.perf_counter() - t0
    return {"value": S * N * steps / el, "unit": "strand-point updates/s", "cores": threads, "kind": "port",
            "sample": "%d strands x %d points of the same workload, %d steps, %.1f s; CPU restatement of compute.comp "
                      "(oracle/oracle.c, OpenMP; lavapipe/Vulkan unavailable in this image)" % (S, N, steps, el),
            "steps": steps, "seconds": el, "strands": S}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import rvh_b200 as rvh  # scenes only
    S_full, N, L, flags_s, desc = WORKLOADS[args.workload]
    # each "step" of this arm is one oracle step over a bounded sample
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    S = min(S_full, 131072)
    cols = rvh.scenes.bench_colliders()
    st = rvh.scenes.synthetic_head(S, N, L)
    rest = np.float32(L) / np.float32(N - 1)
    of = (orc.GRID_ON if "grid" in flags_s else 0) | (orc.WIND_B if "windB" in flags_s else 0)
    p = orc.default_params(S, N, of, rest_length=rest)
    grid = orc.new_grid(p)
    threads = orc.lib().orc_max_threads()
    L_ = orc.lib()
    colp = np.ascontiguousarray(cols, np.float32)

    def one(t):
        L_.orc_step_parallel(C.byref(p), orc._f(colp), DT, t, orc._f(st), grid.ctypes.data_as(orc._i64p), threads)

    for w in range(args.warmup):
        one(DT * w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one(DT * (args.warmup + k))
    el = time.perf_counter() - t0
    val = S * N * args.steps / el
    sample = "%d strands x %d points per step (bounded sample of the workload), %d host threads" % (S, N, threads)
    out = {"impl": "reference", "metric": "strand-point updates/sec", "value": val, "unit": "strand-point updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "description": desc, "reference_arm": "CPU restatement of compute.comp (oracle/oracle.c, OpenMP): the reference is a GLSL compute shader and this image has no Vulkan/lavapipe"},
           "cpu_baseline": {"value": val, "unit": "strand-point updates/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "strand-point updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="ns_full", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spt", type=int, default=0, help="strands per thread (0 auto)")
    ap.add_argument("--flags", default=None, help="override the workload's feature flags, e.g. grid+windB (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import rvh_b200 as rvh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(rvh.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    S, N, L, flags_s, desc = WORKLOADS[args.workload]
    if args.flags is not None:
        flags_s = args.flags
        desc += " [flags overridden: %s]" % flags_s
    flags = parse_flags(rvh, flags_s)
    grid_on = bool(flags & rvh.GRID_ON)
    rest = float(np.float32(L) / np.float32(N - 1))
    cols = rvh.scenes.bench_colliders()

    # synthetic inputs in pinned host memory (global strand ids => every rank makes its own shard)
    aos_bytes = S * 48 * N
    pinned = torch.empty(aos_bytes // 4, dtype=torch.float32, pin_memory=True)
    host = pinned.numpy().reshape(S, 3, N, 4)
    rvh.scenes.synthetic_head(S, N, L, first_strand=rank * S, colliders=cols, out=host)

    cfg = rvh.default_config(S, N, flags=flags, device=local, rest_length=rest, strands_per_thread=args.spt)
    sim = rvh.HairSim(cfg, rank=rank, nranks=world, nccl_id=nccl_id)
    sim.set_colliders(cols)
    sim.upload_ptr(pinned.data_ptr(), aos_bytes)
    sim.sync()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sim.sync()

    # ---- device-resident timed region -------------------------------------------------------
    sim.step_n(max(args.warmup, 1), DT, 0.0, timed=True)
    sim.profile_enable(True)
    sim.profile_read()
    launches0 = sim.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms = sim.step_n(args.steps, DT, DT * args.warmup, timed=True)     # CUDA events on the context's stream
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    prof = sim.profile_read()
    sim.profile_enable(False)
    launches = sim.kernel_launches() - launches0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * S * N * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak = 6650.0; peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    b1, b2 = bytes_per_strand(N, grid_on)
    k1 = prof["ftl_step"]
    k1_ms = k1["ms"] / max(k1["launches"], 1)
    achieved = S * b1 / (k1_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_ftl_step", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": S * b1, "avg_launch_ms": k1_ms,
                "per_kernel_ms": {k: (v["ms"] / v["launches"] if v["launches"] else 0.0) for k, v in prof.items()},
                "step_bytes": S * (b1 + b2), "step_frac": (S * (b1 + b2) / (ms / args.steps * 1e-3) / 1e9) / peak}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get(args.workload, {}).get("k_ftl_step")
        except Exception:
            pass

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------
    e2e = None
    e2e_resident = None
    if not args.no_e2e:
        barrier()
        sim.step_host_ptr(pinned.data_ptr(), aos_bytes, DT, 0.0)            # warm-up
        barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            sim.step_host_ptr(pinned.data_ptr(), aos_bytes, DT, DT * k)      # H2D Strand[S] + step + D2H Strand[S]
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); el = float(t.item())
        e2e = {"value": world * S * N * args.e2e_steps / el, "unit": "strand-point updates/s",
               "h2d_bytes_per_step": aos_bytes + cols.nbytes + 8, "d2h_bytes_per_step": aos_bytes,
               "steps": args.e2e_steps, "what": "rvh_step_host: upload Strand[S] AoS from pinned host memory + step + download Strand[S] AoS, every step"}
        # the reference's own per-frame contract: state stays on the GPU, only Time + Collider UBOs go in (Scene.cpp:78-87,133)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            sim.set_colliders(cols)
            sim.step(DT, DT * k)
            sim.draw_indirect()
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); el = float(t.item())
        e2e_resident = {"value": world * S * N * args.steps / el, "unit": "strand-point updates/s",
                        "h2d_bytes_per_step": cols.nbytes + 8, "d2h_bytes_per_step": 16,
                        "what": "per-frame API as the reference drives it: rvh_set_colliders (UBO) + rvh_step + rvh_draw_indirect read-back; strand state resident"}
    sim.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, flags_s)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    out = {
        "metric": "strand-point updates/sec", "value": value, "unit": "strand-point updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "strands_per_gpu": S, "points_per_strand": N,
                   "dt": DT, "l2": "state %.0f MB per GPU > 126 MB L2, no flush needed" % (S * N * 24 / 1e6) if S * N * 24 > 126e6 else "state %.1f MB is L2-resident (launch/latency-bound config)" % (S * N * 24 / 1e6),
                   "parallelism": "strand-sharded x%d, NCCL int64 grid all-reduce per step" % world if world > 1 else "1 GPU",
                   "strands_per_thread": int(sim.cfg.strands_per_thread)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_resident": e2e_resident,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
