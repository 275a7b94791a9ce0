#!/usr/bin/env python
"""bench.py — ICP iterations/s of the hot path on synthetic Velodyne-like clouds.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--mode auto|shard|pairs]

Workloads (--config; BASELINE.json `configs`):
  c2plane  (default) the north-star target: 1 M-point reading vs 1 M-point reference, KDTreeMatcher knn 1,
           TrimmedDist 0.75, PointToPlaneErrorMinimizer, reference normals from SurfaceNormalDataPointsFilter knn 20
  c2       BASELINE configs[1]: the same with PointToPointErrorMinimizer (no normals)
  c3       configs[2]: 10 M-point map (10 scans, normals knn 20) + 1 M-point reading, point-to-plane
  c4       configs[3]: 2 M x 2 M, knn 10 with maxDist 2 m, MaxDist 1 m x MedianDist 3, PointToPlaneWithCov, normals knn 20
  c5       configs[4]: a batch of independent 200 k x 200 k scan pairs streamed through every GPU (pairs round-robin)
A *step* is one ICP iteration (match -> select / weights -> minimise -> compose -> check).

  value  : iterations/s with reading + reference structure resident in HBM: K iterations per repetition, each
           bracketed by CUDA events on the context's stream with an L2 flush in between, `reps` repetitions (each a
           fresh registration: reset, no matches carried over), the MEDIAN repetition reported.
  e2e    : the same metric for whole registrations through the public API (pm.ICP) from pinned HOST buffers:
           reference upload + structure build + normals + reading upload + K iterations + result download inside
           the timed region, median of `e2e_reps` registrations.
  N > 1  : one process per GPU (torchrun).  Default mode: the queries of ONE registration sharded over the ranks
           against a replicated reference, per-iteration exchanges fused into the kernels over NVLink peer mailboxes
           (comm.cuh), map normals computed per slice + all-gathered -> "scaling": "strong".  --mode pairs: one
           independent scan pair per rank, no collective -> "weak" (also reported under extra at N > 1).
  --impl reference : the CPU restatement of the reference path (oracle/, kind "port": the reference itself cannot
           be built in this image) on all host threads, same config, normals included in its e2e.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "icp_iterations_per_s"
UNIT = "iterations/s"
INF = float("inf")

# filters / minimizer are named here and resolved against capi / the oracle binding where they are used
CONFIGS = {
    "c2plane": dict(nq=1_000_000, nr=1_000_000, map_scans=0, knn=1, max_dist=INF, filters=[("TRIMMEDDIST", 0.75)], minimizer="P2PLANE", normals_knn=20,
                    label="north-star target: 1M x 1M, KDTreeMatcher knn=1, TrimmedDist 0.75, PointToPlane, SurfaceNormal knn=20 on the reference "
                          "(BASELINE configs[1] with the point-to-plane minimiser)"),
    "c2": dict(nq=1_000_000, nr=1_000_000, map_scans=0, knn=1, max_dist=INF, filters=[("TRIMMEDDIST", 0.75)], minimizer="P2POINT", normals_knn=0,
               label="BASELINE configs[1]: 1M x 1M, KDTreeMatcher knn=1, TrimmedDist 0.75, PointToPoint"),
    "c3": dict(nq=1_000_000, nr=10_000_000, map_scans=10, knn=1, max_dist=INF, filters=[("TRIMMEDDIST", 0.75)], minimizer="P2PLANE", normals_knn=20,
               label="BASELINE configs[2]: 10M-pt map (10 scans), SurfaceNormal knn=20, 1M-pt reading, knn=1, TrimmedDist 0.75, PointToPlane"),
    "c4": dict(nq=2_000_000, nr=2_000_000, map_scans=0, knn=10, max_dist=2.0, filters=[("MAXDIST", 1.0), ("MEDIANDIST", 3.0)], minimizer="P2PLANE_COV",
               normals_knn=20, label="BASELINE configs[3]: 2M x 2M, knn=10 maxDist 2 m, MaxDist 1 m x MedianDist 3, PointToPlaneWithCov, SurfaceNormal knn=20"),
    "c5": dict(nq=200_000, nr=200_000, map_scans=0, knn=1, max_dist=INF, filters=[("TRIMMEDDIST", 0.75)], minimizer="P2PLANE", normals_knn=10, pairs=1024,
               label="BASELINE configs[4]: independent 200k x 200k scan pairs, knn=1, TrimmedDist 0.75, PointToPlane, SurfaceNormal knn=10, "
                     "pairs round-robin over the GPUs"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2plane", choices=sorted(CONFIGS))
    ap.add_argument("--points", type=int, default=0, help="override the reading / reference sizes (quick runs)")
    ap.add_argument("--pairs", type=int, default=0, help="c5: number of pairs (default 1024, bounded by --max-seconds)")
    ap.add_argument("--mode", default="auto", choices=["auto", "pairs", "shard"])
    ap.add_argument("--reps", type=int, default=11, help="repetitions of the K-iteration timed loop (median reported)")
    ap.add_argument("--e2e-reps", type=int, default=10, help="whole registrations timed for e2e (median)")
    ap.add_argument("--rings", type=int, default=0, help="scan rings (0: isotropic sampling, 64: SURVEY 8d's 64 x N/64 layout)")
    ap.add_argument("--cpu-sample-iters", type=int, default=3)
    ap.add_argument("--no-extra", action="store_true", help="skip the explanatory extra sections")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    return ap.parse_args()


def resolved(args, name=None):
    cfg = dict(CONFIGS[name or args.config])
    if args.points:
        scale = args.points / cfg["nq"]
        cfg["nq"] = args.points
        cfg["nr"] = max(1000, int(cfg["nr"] * scale))
    if args.pairs and "pairs" in cfg:
        cfg["pairs"] = args.pairs
    return cfg


def config_dict(args, cfg, world, mode):
    return {
        "workload": "synthetic Velodyne-like clouds; " + cfg["label"] + "; %d iterations per registration" % args.steps,
        "name": args.config, "points_reading": cfg["nq"], "points_reference": cfg["nr"], "knn": cfg["knn"],
        "max_dist": None if cfg["max_dist"] == INF else cfg["max_dist"],
        "outlier_filters": ["%s(%g)" % f for f in cfg["filters"]], "minimizer": cfg["minimizer"], "normals_knn": cfg["normals_knn"],
        "iterations": args.steps, "repetitions": args.reps,
        "ring_layout": ("%d rings" % args.rings) if args.rings else "isotropic angular sampling (256 rings x 3906 azimuth steps at 1M; see extra.ring_layout_64)",
        "multi_gpu": {"single": "one GPU", "pairs": "one independent scan pair per rank, no collective",
                      "shard": "queries of ONE registration dealt out to the ranks (round-robin columns), reference replicated; select histograms and normal-equation sums "
                               "exchanged inside the producing kernels over NVLink peer mailboxes; map normals per slice + ncclAllGather"}[mode],
        "l2": "flushed between timed iterations (256 MiB memset)",
        "world_size": world,
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks
    line): an NVML polling thread (~1 kHz); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.proc, self.path, self.thread, self.nvml = None, None, None, None
        self.sm, self.mx, self.bits, self.running = [], None, 0, False

    def _nvml_loop(self):
        nv, h = self.nvml, self.handle
        while self.running:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:
                self.handle = nv.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid else nv.nvmlDeviceGetHandleByIndex(self.index)
            except Exception:
                self.handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.nvml, self.running = nv, True
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
            nv, reasons = self.nvml, []
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if self.bits & int(bit):
                    reasons.append(name)
            if self.sm:
                out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                   "source": "nvidia-smi"}
        return out


def gpu_uuid(torch, index):
    try:
        return "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
    except Exception:
        return None


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return v, t


def host_threads():
    """every core this process may run on — not OMP_NUM_THREADS, which torchrun sets to 1 for each rank it launches"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_clouds(cfg, args, pair_seed=0):
    """(reading, reference, T_gt) of a workload; c3's reference is the 10-scan world map"""
    from libpointmatcher_b200 import synth
    rings = args.rings or None
    if cfg["map_scans"]:
        rf = synth.world_map(cfg["nr"], cfg["map_scans"])
        rd = synth.scan(cfg["nq"], synth.READING_POSE, rings=rings, seed=synth.SEED + 1)
        return rd, rf, synth.READING_POSE
    return synth.scan_pair(cfg["nq"], cfg["nr"], pair_seed=pair_seed, rings=rings)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def reference_install():
    """the reference itself cannot be built in this image (no Eigen3 / Boost headers / libnabo), so the CPU arm is the
    oracle port; should a driver-written install ever appear under baseline/_ref, say so on the line instead of hiding it"""
    for rel in ("baseline/_ref/bin/pmicp", "baseline/_ref/lib/libpointmatcher.so"):
        if os.path.exists(os.path.join(ROOT, rel)):
            return "found %s (not used: the arm times the oracle port, kind = port)" % rel
    return "absent (reference not buildable here; oracle port timed)"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on all host threads (rank 0 only), same config: the reference
    filter (SurfaceNormal) is inside its e2e exactly as it is inside ours."""
    if rank != 0:
        return None
    from oracle import binding as orc
    orc.build()
    cfg = resolved(args)
    threads = host_threads()
    if args.config == "c5":
        return run_reference_pairs(args, cfg, orc, threads, world)
    rd, rf, _ = make_clouds(cfg, args)
    filters = [(getattr(orc, "FILTER_" + n), v) for n, v in cfg["filters"]]
    kw = dict(knn=cfg["knn"], max_dist=cfg["max_dist"], filters=filters, minimizer=getattr(orc, "MIN_" + cfg["minimizer"]), nthreads=threads)
    t0 = time.perf_counter()
    nrm = orc.surface_normals(rf, knn=cfg["normals_knn"], nthreads=threads)["normals"] if cfg["normals_knn"] else None
    t_normals = time.perf_counter() - t0
    if args.warmup > 0:
        orc.icp(rd, rf, nrm, max_iterations=1, **kw)
    t0 = time.perf_counter()
    orc.icp(rd, rf, nrm, max_iterations=args.steps, **kw)
    wall = time.perf_counter() - t0
    tm = orc.last_timings()
    value = tm["iterations"] / tm["loop_s"]
    e2e = tm["iterations"] / (wall + t_normals)
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tm["loop_s"] / max(1, tm["iterations"]), "higher_is_better": True, "scaling": "strong" if world > 1 and args.mode != "pairs" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, cfg, world, "single"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d iterations of the full workload, OpenMP over queries; kd-tree build (%.2f s) and SurfaceNormal knn=%d on the reference "
                                   "(%.2f s) excluded from value, included in e2e" % (tm["iterations"], tm["build_s"], cfg["normals_knn"], t_normals)},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extra": {"match_share": tm["match_s"] / tm["loop_s"], "build_s": tm["build_s"], "normals_s": t_normals, "loop_s": tm["loop_s"],
                  "reference_install": reference_install()},
    }


def run_reference_pairs(args, cfg, orc, threads, world):
    """c5 on the CPU: pairs one after the other, every pair on all threads (OpenMP over queries), bounded sample"""
    filters = [(getattr(orc, "FILTER_" + n), v) for n, v in cfg["filters"]]
    n_pairs = max(1, min(cfg["pairs"], 4))
    t_all, iters = 0.0, 0
    for j in range(n_pairs):
        rd, rf, _ = make_clouds(cfg, args, pair_seed=j + 1)
        t0 = time.perf_counter()
        nrm = orc.surface_normals(rf, knn=cfg["normals_knn"], nthreads=threads)["normals"]
        orc.icp(rd, rf, nrm, knn=cfg["knn"], max_dist=cfg["max_dist"], filters=filters, minimizer=getattr(orc, "MIN_" + cfg["minimizer"]),
                max_iterations=args.steps, nthreads=threads)
        t_all += time.perf_counter() - t0
        iters += orc.last_timings()["iterations"]
    value = iters / t_all
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / max(1, iters), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, world, "pairs"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of the %d pairs, each from host buffers (normals + kd-tree build + loop), one after the other on all threads" % (n_pairs, cfg["pairs"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "extra": {"pairs_per_s": n_pairs / t_all},
    }


# ------------------------------------------------------------------------------------------------
class Timed:
    """the resident-data measurement of one workload on one context"""

    def __init__(self, torch, capi, ctx, dev, dist_on):
        self.torch, self.capi, self.ctx, self.dev, self.dist_on = torch, capi, ctx, dev, dist_on
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier(self):
        if self.dist_on:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def repetition(self, params, steps):
        """one fresh registration of exactly `steps` iterations, each bracketed by events with an L2 flush before it; an
        iteration slot voided by capped matching stays in the total and is topped up.  Returns (ms list, result)."""
        torch, ctx = self.torch, self.ctx
        ctx.icp_reset(None)
        ms, todo = [], steps
        while todo > 0:
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(todo)]
            stops = [torch.cuda.Event(enable_timing=True) for _ in range(todo)]
            for i in range(todo):
                with torch.cuda.stream(self.stream):
                    self.flush.zero_()
                starts[i].record(self.stream)
                ctx.icp_enqueue(params, 1)
                stops[i].record(self.stream)
            ctx.sync()
            res = ctx.icp_result()
            ms += [s.elapsed_time(e) for s, e in zip(starts, stops)]
            todo = steps - res["iterations"]
        return ms, res

    def loop(self, params, steps, warmup, reps):
        """W warm-up iterations, then `reps` repetitions; per-repetition totals are the max over ranks"""
        ctx = self.ctx
        ctx.icp_reset(None)
        if warmup > 0:
            ctx.icp_enqueue(params, warmup)
        ctx.sync()
        ctx.timing_collect()
        totals, first, last, res = [], [], [], None
        self.barrier()
        for _ in range(reps):
            ms, res = self.repetition(params, steps)
            totals.append(float(sum(ms)))
            first.append(ms[0])
            last.append(ms[-1])
        self.barrier()
        t = self.torch.tensor(totals, dtype=self.torch.float64, device=self.dev)
        if self.dist_on:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        totals = [float(x) for x in t.cpu()]
        return {"totals_ms": totals, "median_ms": float(np.median(totals)), "first_iteration_ms": float(np.median(first)),
                "last_iteration_ms": float(np.median(last)), "result": res}


def measure(args, cfg, torch, capi, pm, pmdist, rank, world, local_rank, mode, reps, clouds=None, want_e2e=True, label=""):
    """value / stage times / roofline inputs / e2e of one workload in one multi-GPU mode.  Returns a dict (every rank)."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    sharded = dist_on and mode == "shard"
    rd, rf, T_gt = clouds if clouds is not None else make_clouds(cfg, args, pair_seed=0 if (not dist_on or sharded) else rank)
    if sharded:
        rd_local = pmdist.shard_take(rd, rank, world)
    else:
        rd_local = rd
    rd_pin, _k1 = pinned_copy(rd)          # e2e hands pm.ICP the WHOLE reading; the sharded ICP takes its slice
    rf_pin, _k2 = pinned_copy(rf)
    filters = [(getattr(capi, "FILTER_" + n), v) for n, v in cfg["filters"]]
    minimizer = getattr(capi, "MIN_" + cfg["minimizer"])
    params = capi.make_params(knn=cfg["knn"], max_dist=cfg["max_dist"], filters=filters, minimizer=minimizer, max_iterations=max(args.steps, 1))

    # ---- e2e: whole registrations through the public API from pinned host buffers — measured FIRST, while its context is the
    # only one on the GPU (the library keeps programmatic dependent launch and the in-kernel select for that case) ------------
    e2e_out = {}
    if want_e2e and not args.no_e2e:
        icp = pm.ICP(local_rank)
        icp.matcher = pm.KDTreeMatcher({"knn": str(cfg["knn"]), "maxDist": "inf" if cfg["max_dist"] == INF else repr(cfg["max_dist"])})
        fl = {"TRIMMEDDIST": lambda v: pm.TrimmedDistOutlierFilter({"ratio": repr(v)}), "MAXDIST": lambda v: pm.MaxDistOutlierFilter({"maxDist": repr(v)}),
              "MEDIANDIST": lambda v: pm.MedianDistOutlierFilter({"factor": repr(v)})}
        icp.outlierFilters = pm.OutlierFilters([fl[n](v) for n, v in cfg["filters"]])
        icp.errorMinimizer = {"P2POINT": pm.PointToPointErrorMinimizer, "P2PLANE": pm.PointToPlaneErrorMinimizer,
                              "P2PLANE_COV": pm.PointToPlaneWithCovErrorMinimizer}[cfg["minimizer"]]()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(args.steps)})]
        icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": str(cfg["normals_knn"])})] if cfg["normals_knn"] else []
        if sharded:
            pmdist.init_comm(icp.ctx, capi)
            icp.setSharded(rank, world)
        reading, reference = pm.DataPoints(rd_pin), pm.DataPoints(rf_pin)
        secs, n_it = [], 0
        for i in range(args.e2e_reps + 1):  # the first call is the warm-up (allocations, first-use costs)
            if dist_on:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            T_e2e = icp(reading, reference)
            dt = time.perf_counter() - t0
            n_it = icp.iterationCount
            if i > 0:
                secs.append(dt)
        t = torch.tensor(secs, dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        med = float(np.median(t.cpu().numpy()))
        h2d = (rd_local.nbytes + rf_pin.nbytes)
        e2e_out["e2e"] = {"value": n_it * (world if (dist_on and not sharded) else 1) / med, "unit": UNIT,
                      "h2d_bytes_per_step": h2d / max(1, n_it), "d2h_bytes_per_step": (64.0 + 2 * 3000.0) / max(1, n_it),
                      "seconds_per_registration": med, "registrations_timed": len(secs), "api": "libpointmatcher_b200.pm.ICP (Python mirror over the C ABI)",
                      "note": "whole registrations of %d iterations: H2D of both clouds (pinned) + structure build%s + loop + result D2H, median; "
                              "bytes are per registration and rank divided by iterations" % (n_it, " + SurfaceNormal knn=%d" % cfg["normals_knn"] if cfg["normals_knn"] else "")}
        e2e_out["e2e_T"] = np.asarray(T_e2e, np.float64).tolist()
        icp.ctx.close()

    ctx = capi.Context(local_rank)
    if sharded:
        pmdist.init_comm(ctx, capi)
    # reference centred on its mean, reading moved into that frame — the host bookkeeping of ICP::compute
    # (ICP.cpp:291-299, 345-347), done once outside the timed loop
    mean = pm.sequential_mean(rf_pin)
    rf_c = rf_pin.copy()
    rf_c[:, :3] -= mean[:3]
    T_in = np.eye(4, dtype=np.float32)
    T_in[:3, 3] = -mean[:3]
    ctx.set_reference(rf_c)
    t_normals_ms = None
    if cfg["normals_knn"]:
        ctx.sync()
        t0 = time.perf_counter()
        ctx.ref_compute_normals(knn=cfg["normals_knn"])
        ctx.sync()
        t_normals_ms = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.ref_compute_normals(knn=cfg["normals_knn"])   # second run: allocations and first-use costs behind it
        ctx.sync()
        t_normals_ms = min(t_normals_ms, 1e3 * (time.perf_counter() - t0))
    ctx.set_reading(rd_local)
    ctx.reading_apply_transform(T_in)

    tm = Timed(torch, capi, ctx, dev, dist_on)
    sampler = ClockSampler(local_rank, gpu_uuid(torch, local_rank))
    launches0 = ctx.launch_count
    ctx.timing_enable(True)
    sampler.start()
    loop = tm.loop(params, args.steps, args.warmup, reps)
    clocks = sampler.stop()
    stage = ctx.timing_collect()
    ctx.timing_enable(False)
    launches = ctx.launch_count - launches0
    res = loop["result"]
    units = args.steps * (world if (dist_on and not sharded) else 1)
    out = {
        "mode": mode if dist_on else "single", "value": units / (loop["median_ms"] * 1e-3), "ms_per_step": loop["median_ms"] / args.steps,
        "repetition_totals_ms": loop["totals_ms"], "first_iteration_ms": loop["first_iteration_ms"], "last_iteration_ms": loop["last_iteration_ms"],
        "stage_ms_per_iteration": {k: v[0] / max(1, reps * args.steps) for k, v in stage.items()}, "stage_launches": {k: v[1] for k, v in stage.items()},
        "gpu_launches": launches, "clocks": clocks, "iterations_executed": res["iterations"], "voided_slots": res["cap_redos"],
        "normals_ms": t_normals_ms, "nq_local": len(rd_local), "nr": len(rf),
    }
    T_full = pm.mat4_mul(pm.mat4_mul(np.linalg.inv(T_in.astype(np.float64)).astype(np.float32), res["T_iter"]), T_in)
    out["translation_error_vs_ground_truth_m"] = float(np.linalg.norm(T_full[:3, 3].astype(np.float64) - np.asarray(T_gt)[:3, 3]))

    # back-to-back variant (no flush, one event pair) — how the loop runs in production
    ctx.icp_reset(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm.barrier()
    e0.record(tm.stream)
    ctx.icp_enqueue(params, args.steps)
    e1.record(tm.stream)
    ctx.sync()
    b2b_it = max(1, ctx.icp_result()["iterations"])
    out["back_to_back_ms_per_step"] = pmdist.max_over_ranks(e0.elapsed_time(e1), dev) / b2b_it
    out.update(e2e_out)
    out["ctx"], out["tm"], out["params"], out["T_in"], out["rd"], out["rf"], out["rf_c"], out["rd_pin"], out["rf_pin"], out["T_gt"] = ctx, tm, params, T_in, rd, rf, rf_c, rd_pin, rf_pin, T_gt

    return out


def cpp_e2e(args, cfg, rd, rf):
    """the same registrations through the C++ host mirror (tools/host_e2e.cpp -> PointMatcher<float>::ICP -> C ABI): the host side
    the north star names.  Returns the e2e dict or None (no compiler, c5, ...)."""
    import subprocess
    import tempfile
    name = args.config if args.config in ("c2plane", "c2", "c3", "c4") else None
    if name is None:
        return None
    try:
        host = os.path.join(ROOT, "libpointmatcher_b200", "host")
        libdir = os.path.join(ROOT, "libpointmatcher_b200")
        src = os.path.join(ROOT, "tools", "host_e2e.cpp")
        exe = os.path.join(ROOT, "tools", "_bin", "host_e2e")
        deps = [src, os.path.join(libdir, "libpmgpu.so"), os.path.join(ROOT, "include", "pmgpu.h")] + [os.path.join(host, f) for f in os.listdir(host)]
        if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", host, "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-L", libdir, "-lpmgpu",
                                   "-Wl,-rpath," + libdir], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp:
            a, b = os.path.join(tmp, "reading.f32"), os.path.join(tmp, "reference.f32")
            np.ascontiguousarray(rd, np.float32).tofile(a)
            np.ascontiguousarray(rf, np.float32).tofile(b)
            r = subprocess.run([exe, a, b, str(len(rd)), str(len(rf)), name, str(args.steps), str(args.e2e_reps)], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True, timeout=600)
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        d = json.loads(r.stdout.strip().splitlines()[-1])
        n_it = d["iterations"]
        return {"value": n_it / d["seconds_per_registration"], "unit": UNIT, "h2d_bytes_per_step": (rd.nbytes + rf.nbytes) / max(1, n_it),
                "d2h_bytes_per_step": (64.0 + 2 * 3000.0) / max(1, n_it), "seconds_per_registration": d["seconds_per_registration"],
                "registrations_timed": d["registrations_timed"], "api": "PointMatcher<float>::ICP of the C++ host mirror (tools/host_e2e.cpp) over the C ABI",
                "note": "whole registrations of %d iterations from page-locked host matrices: H2D of both clouds + structure build%s + loop + result D2H, median"
                        % (n_it, " + SurfaceNormal knn=%d" % cfg["normals_knn"] if cfg["normals_knn"] else ""), "T": d["T"]}
    except Exception as e:  # the Python mirror's figure stands
        return {"error": str(e)[-300:]}


def strip(m):
    return {k: v for k, v in m.items() if k not in ("ctx", "tm", "params", "T_in", "rd", "rf", "rf_c", "rd_pin", "rf_pin", "T_gt", "e2e_T")}


def knn_throughput(args, m, capi, orc_threads):
    """exact staged KDTreeMatcher::findClosests throughput (pmgpu_knn, nothing capped) for k = 1, 10, 20: a cold search
    (no previous matches) and a seeded one (the previous matches of the same reading under a slightly different
    transform), queries/s with ids + dists left on the device; the CPU port beside them."""
    ctx = m["ctx"]
    out = {}
    T0 = np.eye(4, dtype=np.float32)
    T1 = np.eye(4, dtype=np.float32)
    T1[:3, 3] = (0.01, -0.01, 0.005)
    nq = ctx.nq
    for k in (1, 10, 20):
        ctx.set_reading(m["rd"])
        ctx.reading_apply_transform(m["T_in"])
        ctx.timing_collect()
        ctx.timing_enable(True)
        ctx.knn(T0, k, 0.0, np.inf, download=False)
        cold = ctx.timing_collect()["knn"][0]
        ctx.knn(T1, k, 0.0, np.inf, download=False)
        seeded = ctx.timing_collect()["knn"][0]
        ctx.timing_enable(False)
        out["k%d" % k] = {"cold_ms": cold, "cold_queries_per_s": nq / (cold * 1e-3), "seeded_ms": seeded, "seeded_queries_per_s": nq / (seeded * 1e-3)}
    if orc_threads:
        from oracle import binding as orc
        tree = orc.KdTree(m["rf_c"])
        q = m["rd"].copy()
        q[:, :3] += m["T_in"][:3, 3]
        sample = q[: min(len(q), 200_000)]
        for k in (1, 10, 20):
            t0 = time.perf_counter()
            tree.knn(sample, k, nthreads=orc_threads)
            dt = time.perf_counter() - t0
            out["k%d" % k]["cpu_queries_per_s"] = len(sample) / dt
        out["cpu_note"] = "oracle kd-tree (leaf 8, sorted linear heap), OpenMP over queries, %d threads, first %d queries" % (orc_threads, len(sample))
    return out


def run_pairs_stream(args, cfg, torch, capi, pm, pmdist, rank, world, local_rank):
    """c5: a batch of independent registrations streamed through this rank's GPU: pairs rank, rank + G, ... (round-robin,
    evaluations/eval_solution.cpp:250-271); `streams` contexts per GPU so that pair j + 1's upload, structure build and
    normals overlap pair j's iteration loop."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    n_pairs = cfg["pairs"]
    mine = pmdist.pair_assignment(n_pairs, rank, world)
    distinct = min(8, len(mine)) or 1   # distinct synthetic pairs generated per rank; the stream cycles through them
    clouds = [make_clouds(cfg, args, pair_seed=1 + rank + world * j) for j in range(distinct)]
    pinned = [(pinned_copy(rd), pinned_copy(rf)) for rd, rf, _ in clouds]
    streams = 3
    icps = []
    for s in range(streams):
        icp = pm.ICP(local_rank)
        icp.matcher = pm.KDTreeMatcher({"knn": str(cfg["knn"])})
        icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": repr(cfg["filters"][0][1])})])
        icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(args.steps)})]
        icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": str(cfg["normals_knn"])})]
        icps.append(icp)

    def work(s, todo, out):
        torch.cuda.set_device(local_rank)
        for j in todo:
            (rd, _a), (rf, _b) = pinned[j % distinct]
            T = icps[s](pm.DataPoints(rd), pm.DataPoints(rf))
            out.append((j, icps[s].iterationCount, T))

    def run(todo_all):
        outs = [[] for _ in range(streams)]
        th = [threading.Thread(target=work, args=(s, todo_all[s::streams], outs[s])) for s in range(streams)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, [x for o in outs for x in o]

    run(list(range(min(len(mine), 2 * streams))))  # warm-up
    sampler = ClockSampler(local_rank, gpu_uuid(torch, local_rank))
    launches0 = sum(i.ctx.launch_count for i in icps)
    if dist_on:
        dist.barrier()
    sampler.start()
    dt, results = run(list(range(len(mine))))
    clocks = sampler.stop()
    launches = sum(i.ctx.launch_count for i in icps) - launches0
    dt_max = pmdist.max_over_ranks(dt, dev)
    iters = sum(r[1] for r in results)
    t = torch.tensor([float(iters), float(len(results))], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t)
    iters_all, pairs_all = float(t[0].item()), float(t[1].item())
    # accuracy of the stream: every distinct pair's transform against its ground truth
    errs = []
    for j, _, T in results[: distinct]:
        T_gt = np.asarray(clouds[j % distinct][2])
        errs.append(float(np.linalg.norm(np.asarray(T, np.float64)[:3, 3] - T_gt[:3, 3])))
    for i in icps:
        i.ctx.close()
    if rank != 0:
        return None
    value = iters_all / dt_max
    nbytes = sum(p[0][0].nbytes + p[1][0].nbytes for p in pinned) / distinct
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / max(1.0, iters_all / world),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, world, "pairs"), "clocks": clocks,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": nbytes / max(1, args.steps), "d2h_bytes_per_step": (64.0 + 6000.0) / max(1, args.steps),
                "note": "every pair comes from pinned host buffers: `value` IS the end-to-end figure for this workload"},
        "gpu_launches": launches,
        "roofline": None, "cpu_baseline": None,
        "extra": {"pairs_per_s": pairs_all / dt_max, "pairs": int(pairs_all), "seconds": dt_max, "contexts_per_gpu": streams,
                  "distinct_pairs_per_rank": distinct, "translation_error_vs_ground_truth_m_max": max(errs) if errs else None,
                  "note": "total work (the batch) is fixed as N grows -> strong scaling of the batch; no data-path collective"},
    }


def run_ours(args, rank, world, local_rank):
    import torch
    from libpointmatcher_b200 import capi, pm
    from libpointmatcher_b200 import dist as pmdist

    torch.cuda.set_device(local_rank)
    cfg = resolved(args)
    dist_on = world > 1
    if args.config == "c5":
        return run_pairs_stream(args, cfg, torch, capi, pm, pmdist, rank, world, local_rank)
    mode = "single" if not dist_on else ("shard" if args.mode in ("auto", "shard") else "pairs")
    m = measure(args, cfg, torch, capi, pm, pmdist, rank, world, local_rank, mode, args.reps)
    ctx = m["ctx"]
    pk = peaks()
    peak = float(pk.get("hbm_gbs", 6650.0))

    # roofline of the dominant kernel (K2, transform + exact kNN): algorithmic bytes / measured launch time
    knn_ms = m["stage_ms_per_iteration"]["knn"]
    k = cfg["knn"]
    alg_bytes = 16 * m["nq_local"] + 16 * m["nr"] + 8 * k * m["nq_local"]
    achieved = alg_bytes / (knn_ms * 1e-3) / 1e9 if knn_ms > 0 else 0.0
    # dram__bytes_read + dram__bytes_write of K2 per launch (stage 1 + stage 2) from the ncu --set full capture of this workload
    # (profiles/r2_end_knn_c2plane_summary.txt: 51.8 + 0.6 MB and 0.2 MB); null for the workloads that were not captured
    traffic = {("c2plane", 1_000_000): 52.6e6, ("c2", 1_000_000): 52.6e6}.get((args.config, m["nq_local"]))
    roofline = {"bound": "hbm", "kernel": "knn_kernel<%d> + knn_overflow_kernel<%d> (K2: transform + exact nearest neighbours, stage 1 + 2)" % (k, k),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if pk else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": knn_ms,
                "note": "K2 is issue / latency bound (divergent tree search), neither HBM nor FP32 bound; see DESIGN.md K2"}
    extra = {"stage_ms_per_iteration": m["stage_ms_per_iteration"], "back_to_back_ms_per_step": m["back_to_back_ms_per_step"],
             "back_to_back_iterations_per_s": 1e3 / m["back_to_back_ms_per_step"], "first_iteration_ms": m["first_iteration_ms"],
             "last_iteration_ms": m["last_iteration_ms"], "repetition_totals_ms": m["repetition_totals_ms"], "normals_ms_resident": m["normals_ms"],
             "translation_error_vs_ground_truth_m": m["translation_error_vs_ground_truth_m"],
             "capped_matching": {"enabled": os.environ.get("PMGPU_NO_CAP") is None, "voided_slots": m["voided_slots"]}}
    nq_l = m["nq_local"]
    sel_ms, min_ms = m["stage_ms_per_iteration"]["select"], m["stage_ms_per_iteration"]["minimize"]
    per_match = {"P2POINT": 40, "P2PLANE": 56, "P2PLANE_COV": 56}[cfg["minimizer"]]
    hbm = {}
    fused = sel_ms == 0.0 and cfg["filters"]   # the select runs inside the minimiser kernel (select_accumulate_kernel): one timing slot
    stages = (("select_minimize", (4 + per_match) * k * nq_l, min_ms),) if fused else (("select", 3 * 4 * k * nq_l, sel_ms), ("minimize", per_match * k * nq_l, min_ms))
    for name, nbytes, t_ms in stages:
        if t_ms > 0:
            gbs = nbytes / (t_ms * 1e-3) / 1e9
            hbm[name] = {"algorithmic_bytes_per_iteration": nbytes, "ms_per_iteration": t_ms, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}
    if fused:
        hbm["select_minimize"]["note"] = ("one cooperative kernel: exact quantile select (one window pass over the distances when the limit moved "
                                          "as extrapolated, else 2-4) + grid barrier + fused residual/normal-equation sums + solve + compose + checkers")
    extra["hbm_stage_rooflines"] = hbm

    if not args.no_extra and not dist_on:
        # the other axis north_star names: FP32-pipe utilisation of the distance evaluation (8 single-rounded ops each)
        try:
            ctx.timing_collect()
            ctx.timing_enable(True)
            _, _, visits = ctx.knn(m["tm"].ctx.icp_result()["T_iter"], k, 0.0, cfg["max_dist"], download=False)
            t_knn = ctx.timing_collect()["knn"][0]
            ctx.timing_enable(False)
            sm_mhz = (m["clocks"].get("sm_mhz") or 1965.0)
            peak_tops = 148 * 128 * sm_mhz * 1e6 / 1e12
            roofline["fp32"] = {"distance_evaluations_per_launch": visits, "ops_per_evaluation": 8, "launch_ms": t_knn,
                                "achieved_tops": visits * 8 / (t_knn * 1e-3) / 1e12, "peak_tops": peak_tops, "frac": visits * 8 / (t_knn * 1e-3) / 1e12 / peak_tops,
                                "note": "staged, uncapped, exact match at the final transform (every query searched to its true neighbours)"}
        except Exception as e:  # never let the explanatory figure break the bench line
            roofline["fp32"] = {"error": str(e)}
        if args.config in ("c2plane", "c2"):
            try:
                extra["knn_queries_per_s"] = knn_throughput(args, m, capi, 0 if args.no_cpu else host_threads())
            except Exception as e:
                extra["knn_queries_per_s"] = {"error": str(e)}

    ctx.close()   # the extra workloads below bring their own context: each is measured alone on its GPU, like the headline
    # ---- cpu_baseline on this box's host cores (rank 0, N = 1 only) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import binding as orc
        orc.build()
        threads = host_threads()
        filters = [(getattr(orc, "FILTER_" + n), v) for n, v in cfg["filters"]]
        kw = dict(knn=cfg["knn"], max_dist=cfg["max_dist"], filters=filters, minimizer=getattr(orc, "MIN_" + cfg["minimizer"]))
        t0 = time.perf_counter()
        nrm = orc.surface_normals(m["rf"], knn=cfg["normals_knn"], nthreads=threads)["normals"] if cfg["normals_knn"] else None
        t_nrm = time.perf_counter() - t0
        n_it = max(1, args.cpu_sample_iters)
        orc.icp(m["rd"], m["rf"], nrm, max_iterations=n_it, nthreads=threads, **kw)
        t_cpu = orc.last_timings()
        cpu = {"value": t_cpu["iterations"] / t_cpu["loop_s"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d iterations of the same workload (oracle port, OpenMP over queries); kd-tree build %.2f s and SurfaceNormal knn=%d "
                         "%.2f s excluded" % (t_cpu["iterations"], t_cpu["build_s"], cfg["normals_knn"], t_nrm),
               "normals_s": t_nrm, "build_s": t_cpu["build_s"]}
        if cfg["nq"] <= 1_000_000 and cfg["nr"] <= 1_000_000:
            orc.icp(m["rd"], m["rf"], nrm, max_iterations=1, nthreads=1, **kw)
            t1 = orc.last_timings()
            cpu["single_thread_value"] = t1["iterations"] / t1["loop_s"]

    # ---- extra sections -------------------------------------------------------------------------
    if not args.no_extra:
        if args.config == "c2plane" and not args.rings:
            # SURVEY 8d's literal ring layout (64 rings x N / 64 azimuth steps), same workload
            try:
                a64 = argparse.Namespace(**vars(args))
                a64.rings = 64
                m64 = measure(a64, cfg, torch, capi, pm, pmdist, rank, world, local_rank, mode, max(3, args.reps // 3), want_e2e=False)
                m64["ctx"].close()
                extra["ring_layout_64"] = {"value": m64["value"], "ms_per_step": m64["ms_per_step"], "stage_ms_per_iteration": m64["stage_ms_per_iteration"],
                                           "translation_error_vs_ground_truth_m": m64["translation_error_vs_ground_truth_m"],
                                           "note": "64 x 15625: 1 M points on 64 thin circles; 2 cm range noise over ~4 mm point spacing along a ring makes "
                                                   "knn-20 normals ill-defined, which is why the headline samples isotropically"}
            except Exception as e:
                extra["ring_layout_64"] = {"error": str(e)}
        if args.config == "c2plane":
            # BASELINE configs[1] as written (point-to-point), abridged
            try:
                c2 = resolved(args, "c2")
                m2 = measure(args, c2, torch, capi, pm, pmdist, rank, world, local_rank, mode, max(3, args.reps // 3), clouds=(m["rd"], m["rf"], m["T_gt"]),
                             want_e2e=not dist_on)
                m2["ctx"].close()
                extra["c2_point_to_point"] = {"value": m2["value"], "ms_per_step": m2["ms_per_step"], "stage_ms_per_iteration": m2["stage_ms_per_iteration"],
                                              "e2e": m2.get("e2e")}
            except Exception as e:
                extra["c2_point_to_point"] = {"error": str(e)}
        if dist_on and mode == "shard":
            # (a) the replica figure: one independent pair per rank, no collective
            try:
                mp = measure(args, cfg, torch, capi, pm, pmdist, rank, world, local_rank, "pairs", max(3, args.reps // 3))
                mp["ctx"].close()
                extra["pairs_replicas"] = {"value": mp["value"], "ms_per_step": mp["ms_per_step"], "e2e": mp.get("e2e"), "scaling": "weak",
                                           "note": "one independent scan pair per rank, no collective"}
            except Exception as e:
                extra["pairs_replicas"] = {"error": str(e)}
        if args.config == "c2plane" and os.environ.get("PMGPU_BENCH_NO_C4") is None:
            # (b) where sharding pays: configs[3] (2 M, knn 10), the same sharded mode — at N = 1 the single-GPU figure,
            # so that strong-scaling efficiency can be read from the N = 1, 2, 4, 8 lines of one scaling run
            try:
                c4 = resolved(args, "c4")
                m4 = measure(args, c4, torch, capi, pm, pmdist, rank, world, local_rank, mode, 3, want_e2e=False)
                m4["ctx"].close()
                extra["c4_strong_scaling"] = {"value": m4["value"], "ms_per_step": m4["ms_per_step"], "stage_ms_per_iteration": m4["stage_ms_per_iteration"],
                                              "normals_ms_resident": m4["normals_ms"], "n_gpus": world, "scaling": "strong",
                                              "workload": c4["label"]}
            except Exception as e:
                extra["c4_strong_scaling"] = {"error": str(e)}

    # e2e through the C++ host mirror (the north star's host side); the headline e2e is the faster of the two mirrors
    e2e = m.get("e2e")
    if world == 1 and not args.no_e2e:
        ctx.close()   # its buffers go back to the pool before the other process allocates its own
        c = cpp_e2e(args, cfg, m["rd"], m["rf"])
        if c is not None and "error" not in c and e2e is not None:
            T_py = np.asarray(m.get("e2e_T", np.eye(4)), np.float64)
            c["max_abs_diff_vs_python_mirror_T"] = float(np.abs(np.asarray(c.pop("T"), np.float64).reshape(4, 4).T - T_py).max()) if "e2e_T" in m else None
            if c["value"] > e2e["value"]:
                extra["e2e_python_mirror"] = e2e
                e2e = c
            else:
                extra["e2e_cpp_mirror"] = c
        elif c is not None:
            extra["e2e_cpp_mirror"] = c
    ctx.close()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "weak" if mode == "pairs" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, world, mode), "clocks": m["clocks"], "e2e": e2e, "gpu_launches": m["gpu_launches"],
        "roofline": roofline, "cpu_baseline": cpu, "extra": extra, "iterations_executed": m["iterations_executed"],
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the one JSON line and nothing else: whatever a library writes on fd 1 meanwhile (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) is sent to stderr
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), file=out, flush=True)
        return 0
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=out, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
