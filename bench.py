#!/usr/bin/env python
"""bench.py — ICP iterations/s of the hot path on synthetic Velodyne-like clouds.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P] [--mode pairs|shard]

Workload (BASELINE.json configs[1]): 1 M-point reading vs 1 M-point reference, KDTreeMatcher
knn = 1, TrimmedDistOutlierFilter ratio = 0.75, PointToPointErrorMinimizer; a *step* is one ICP
iteration (match -> select/weights -> minimise -> compose).  The same loop with
PointToPlaneErrorMinimizer (the north-star target) is reported under "extra".

  value  : iterations/s with reading + reference structure resident in HBM, timed per iteration
           with CUDA events on the context's stream, L2 flushed between timed iterations.
  e2e    : iterations/s of a whole registration through the public API (pm.ICP) from pinned HOST
           buffers: reference upload + structure build + reading upload + K iterations + result
           download, all inside the timed region.
  N > 1  : one process per GPU (torchrun), one independent scan pair per rank (batched
           align_sequence-style registration, no data-path collective) -> "scaling": "weak";
           --mode shard instead splits the queries of ONE registration over the ranks with the
           NCCL all-reduces of comm.cu (strong scaling), reported for information.
  --impl reference : the CPU restatement of the reference path (oracle/, kind "port": the
           reference itself cannot be built in this image) on all host threads, same config.
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
RATIO = 0.75


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--mode", default="pairs", choices=["pairs", "shard"])
    ap.add_argument("--cpu-sample-iters", type=int, default=3)
    ap.add_argument("--no-extra", action="store_true", help="skip the point-to-plane extra section")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    return ap.parse_args()


def config_dict(args, world):
    return {
        "workload": "synthetic Velodyne-like %d-pt reading vs %d-pt reference, KDTreeMatcher knn=1, TrimmedDist ratio=%.2f, "
                    "PointToPoint, %d iterations (BASELINE configs[1])" % (args.points, args.points, RATIO, args.steps),
        "points_reading": args.points, "points_reference": args.points, "knn": 1, "outlier_filter": "TrimmedDist(%.2f)" % RATIO,
        "minimizer": "PointToPoint", "iterations": args.steps,
        "multi_gpu": ("one independent scan pair per rank, no collective" if args.mode == "pairs"
                      else "queries of one registration sharded over ranks, NCCL all-reduce of histograms + normal equations"),
        "l2": "flushed between timed iterations (256 MiB memset); working set ~62 MB < 126 MB L2",
        "world_size": world,
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks
    line).  The timed region is only tens of milliseconds, so the sampler is an NVML polling thread
    (~1 kHz); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
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


def make_params(capi, minimizer, steps):
    return capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, RATIO)], minimizer=minimizer, max_iterations=max(steps, 1))


# ------------------------------------------------------------------------------------------------
def host_threads():
    """every core this process may run on — not OMP_NUM_THREADS, which torchrun sets to 1 for each rank it launches"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on all host threads (rank 0 only)."""
    if rank != 0:
        return None
    from libpointmatcher_b200 import synth
    from oracle import binding as orc
    orc.build()
    threads = host_threads()
    rd, rf, _ = synth.scan_pair(args.points)
    kw = dict(filters=[(orc.FILTER_TRIMMEDDIST, RATIO)], minimizer=orc.MIN_P2POINT, nthreads=threads)
    if args.warmup > 0:
        orc.icp(rd, rf, max_iterations=min(args.warmup, 1), **kw)
    t0 = time.perf_counter()
    orc.icp(rd, rf, max_iterations=args.steps, **kw)
    wall = time.perf_counter() - t0
    tm = orc.last_timings()
    value = tm["iterations"] / tm["loop_s"]
    e2e = tm["iterations"] / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tm["loop_s"] / max(1, tm["iterations"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d iterations of the full %d x %d workload, OpenMP over queries, kd-tree build excluded from value "
                                   "(%.2f s) and included in e2e" % (tm["iterations"], args.points, args.points, tm["build_s"])},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extra": {"match_share": tm["match_s"] / tm["loop_s"], "build_s": tm["build_s"]},
    }
    return line


# ------------------------------------------------------------------------------------------------
def timed_iterations(ctx, capi, torch, params, steps, warmup, flush_buf, T0=None):
    """W untimed warm-up iterations, reset, then exactly `steps` iterations, each bracketed by
    CUDA events on the context's stream with an L2 flush in between.  An iteration slot whose
    capped match was void (pmgpu.h "capped matching") does not advance the loop: its time stays in
    the total and further timed slots are enqueued until `steps` iterations have executed.
    Returns (ms list, result)."""
    stream = torch.cuda.ExternalStream(ctx.stream)
    ctx.icp_reset(T0)
    if warmup > 0:
        ctx.icp_enqueue(params, warmup)
    ctx.sync()
    ctx.timing_collect()  # drop the warm-up intervals (they contain the lazy module load of the first launch)
    ctx.icp_reset(T0)
    ms, todo = [], steps
    while todo > 0:
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(todo)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(todo)]
        for i in range(todo):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            starts[i].record(stream)
            ctx.icp_enqueue(params, 1)
            stops[i].record(stream)
        ctx.sync()
        res = ctx.icp_result()
        ms += [s.elapsed_time(e) for s, e in zip(starts, stops)]
        todo = steps - res["iterations"]
    return ms, res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from libpointmatcher_b200 import capi, pm, synth
    from libpointmatcher_b200 import dist as pmdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    sharded = dist_on and args.mode == "shard"
    pair_seed = 0 if (not dist_on or sharded) else rank
    rd, rf, T_gt = synth.scan_pair(args.points, pair_seed=pair_seed)
    if sharded:  # contiguous column range of the reading per rank
        lo, hi = pmdist.shard_range(len(rd), rank, world)
        rd_local = np.ascontiguousarray(rd[lo:hi])
    else:
        rd_local = rd
    rd_pin, _k1 = pinned_copy(rd_local)
    rf_pin, _k2 = pinned_copy(rf)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ctx = capi.Context(local_rank)
    if sharded:
        pmdist.init_comm(ctx, capi)

    # reference centred on its mean, reading moved into that frame — the host bookkeeping of
    # ICP::compute (ICP.cpp:291-299, 345-347), done once outside the timed loop
    mean = pm.sequential_mean(rf_pin)
    rf_c = rf_pin.copy()
    rf_c[:, :3] -= mean[:3]
    T_in = np.eye(4, dtype=np.float32)
    T_in[:3, 3] = -mean[:3]
    ctx.set_reference(rf_c)
    ctx.set_reading(rd_pin)
    ctx.reading_apply_transform(T_in)

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline: point-to-point loop, resident data --------------------------------------
    params = make_params(capi, capi.MIN_P2POINT, args.steps)
    sampler = ClockSampler(local_rank, gpu_uuid(torch, local_rank))
    launches0 = ctx.launch_count
    ctx.timing_enable(True)
    barrier()
    sampler.start()
    ms, res = timed_iterations(ctx, capi, torch, params, args.steps, args.warmup, flush_buf)
    barrier()
    clocks = sampler.stop()
    stage = ctx.timing_collect()
    ctx.timing_enable(False)
    launches = ctx.launch_count - launches0
    total_ms = float(sum(ms))
    total_ms_max = pmdist.max_over_ranks(total_ms, dev)
    units = args.steps * (1 if (not dist_on or sharded) else world)
    value = units / (total_ms_max * 1e-3)

    # back-to-back variant (no flush, one event pair) — how the loop runs in production
    stream = torch.cuda.ExternalStream(ctx.stream)
    ctx.icp_reset(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    ctx.icp_enqueue(params, args.steps)
    e1.record(stream)
    ctx.sync()
    b2b_ms = e0.elapsed_time(e1)
    b2b_it = max(1, ctx.icp_result()["iterations"])

    # roofline of the dominant kernel (kNN match): algorithmic bytes / measured launch time
    knn_ms, knn_n = stage["knn"]
    nq_local, nr = len(rd_local), len(rf)
    alg_bytes = 16 * nq_local + 16 * nr + 8 * 1 * nq_local
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    knn_avg_ms = knn_ms / max(1, knn_n)
    achieved = alg_bytes / (knn_avg_ms * 1e-3) / 1e9 if knn_avg_ms > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of one knn_kernel<1> launch from the committed
    # `ncu --set full` capture (profiles/r1_end_summary.txt; cold L2, 1 M x 1 M): 51.6 MB read + 0.7 MB written
    ncu_traffic = 52.2e6 if (nq_local == 1_000_000 and nr == 1_000_000) else None
    roofline = {"bound": "hbm", "kernel": "knn_kernel<1> + knn_overflow_kernel<1> (K2: transform + exact nearest neighbour, stage 1 + stage 2)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": knn_avg_ms,
                "note": "K2 is issue/latency bound (divergent tree search), neither HBM nor FP32 bound; see DESIGN.md K2"}
    # the other axis BASELINE's north_star names: FP32-pipe utilisation of the distance evaluation.  One aligned staged match
    # (outside every timed region) returns the number of reference points examined; 8 single-rounded ops per evaluation
    # (3 sub, 3 mul, 2 add - no FMA, by the bit-exactness rule) against 148 SMs x 128 lanes x SM clock.
    try:
        ctx.timing_collect()
        ctx.timing_enable(True)
        _, _, visits = ctx.knn(res["T_iter"], 1, 0.0, np.inf, download=False)
        t_knn = ctx.timing_collect()["knn"][0]
        ctx.timing_enable(False)
        sm_mhz = (clocks.get("sm_mhz") or 1965.0)
        peak_tops = 148 * 128 * sm_mhz * 1e6 / 1e12
        roofline["fp32"] = {"distance_evaluations_per_launch": visits, "ops_per_evaluation": 8, "launch_ms": t_knn,
                            "achieved_tops": visits * 8 / (t_knn * 1e-3) / 1e12, "peak_tops": peak_tops,
                            "frac": visits * 8 / (t_knn * 1e-3) / 1e12 / peak_tops,
                            "note": "staged, uncapped match at the final transform (all queries searched to their true neighbour)"}
    except Exception as e:  # never let the explanatory figure break the bench line
        roofline["fp32"] = {"error": str(e)}

    # the HBM-bound stages next to it (select = one 4 B read per match and pass, three passes; point-to-point
    # normal equations = 40 B per match), same peak, from the same per-stage CUDA-event timings
    sel_ms, min_ms = stage["select"][0] / max(1, args.steps), stage["minimize"][0] / max(1, args.steps)
    hbm_stages = {}
    for name, nbytes, t_ms in (("select (3 x hist_kernel)", 3 * 4 * nq_local, sel_ms), ("minimize (accumulate_kernel<0>)", 40 * nq_local, min_ms)):
        if t_ms > 0:
            gbs = nbytes / (t_ms * 1e-3) / 1e9
            hbm_stages[name] = {"algorithmic_bytes_per_iteration": nbytes, "ms_per_iteration": t_ms, "achieved": gbs, "peak": peak, "unit": "GB/s",
                                "frac": gbs / peak}
    extra = {
        "hbm_stage_rooflines": hbm_stages,
        "stage_ms_per_iteration": {k: v[0] / max(1, args.steps) for k, v in stage.items()},
        "knn_queries_per_s": (nq_local * (world if dist_on else 1)) / (knn_avg_ms * 1e-3) if knn_avg_ms > 0 else None,
        "back_to_back_iterations_per_s": b2b_it / (b2b_ms * 1e-3),
        "back_to_back_ms_per_step": b2b_ms / b2b_it,
        "per_iteration_ms_first_last": [ms[0], ms[-1]],
        "capped_matching": {"enabled": os.environ.get("PMGPU_NO_CAP") is None, "voided_slots": res["cap_redos"], "timed_slots": len(ms),
                            "note": "fused loop only: the matcher stops at 1.5x the largest squared distance the previous iteration's "
                                    "outlier filters needed; verified every iteration, T bit-identical to the uncapped loop (tests)"},
    }
    if not args.no_extra and not sharded and os.environ.get("PMGPU_NO_CAP") is None:
        # the same loop with the adaptive search radius switched off (every query searched to its true neighbour)
        os.environ["PMGPU_NO_CAP"] = "1"
        ctx_u = capi.Context(local_rank)
        del os.environ["PMGPU_NO_CAP"]
        ctx_u.set_reference(rf_c)
        ctx_u.set_reading(rd_pin)
        ctx_u.reading_apply_transform(T_in)
        ms_u, res_u = timed_iterations(ctx_u, capi, torch, params, args.steps, args.warmup, flush_buf)
        ctx_u.close()
        extra["capped_matching"]["uncapped_iterations_per_s"] = args.steps / (sum(ms_u) * 1e-3)
        extra["capped_matching"]["T_iter_identical_to_uncapped"] = bool((res_u["T_iter"].view(np.uint32) == res["T_iter"].view(np.uint32)).all())

    # ---- e2e: whole registration through the public API from pinned host buffers -----------
    def e2e_once(minimizer_cls, normals):
        icp = pm.ICP(local_rank)
        icp.matcher = pm.KDTreeMatcher({"knn": "1"})
        icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": repr(RATIO)})])
        icp.errorMinimizer = minimizer_cls()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(args.steps)})]
        icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "20"})] if normals else []
        reading, reference = pm.DataPoints(rd_pin), pm.DataPoints(rf_pin)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T = icp(reading, reference)
        dt = time.perf_counter() - t0
        n_it = icp.iterationCount
        icp.ctx.close()
        return T, n_it, dt

    e2e = None
    if not sharded and not args.no_e2e:
        e2e_once(pm.PointToPointErrorMinimizer, False)  # warm-up (allocations, first-use costs)
        barrier()
        T_e2e, n_it, dt = e2e_once(pm.PointToPointErrorMinimizer, False)
        dt_max = pmdist.max_over_ranks(dt, dev)
        e2e = {"value": n_it * (world if dist_on else 1) / dt_max, "unit": UNIT,
               "h2d_bytes_per_step": (rd_pin.nbytes + rf_pin.nbytes) / max(1, n_it), "d2h_bytes_per_step": 64.0 / max(1, n_it) + 0.0,
               "note": "one whole registration of %d iterations per call: H2D of both clouds (pinned) + structure build + loop + 4x4 D2H; "
                       "bytes are per registration divided by iterations" % n_it, "seconds_per_registration": dt_max}

    # ---- extra: the north-star target config (point-to-plane) ------------------------------
    if not args.no_extra and not sharded:
        ctx.ref_compute_normals(knn=20)
        pp = make_params(capi, capi.MIN_P2PLANE, args.steps)
        ctx.timing_enable(True)
        ms_pl, res_pl = timed_iterations(ctx, capi, torch, pp, args.steps, args.warmup, flush_buf)
        st_pl = ctx.timing_collect()
        ctx.timing_enable(False)
        T_full = pm.mat4_mul(pm.mat4_mul(np.linalg.inv(T_in.astype(np.float64)).astype(np.float32), res_pl["T_iter"]), T_in)
        extra["point_to_plane"] = {
            "iterations_per_s": args.steps / (sum(ms_pl) * 1e-3), "ms_per_step": sum(ms_pl) / args.steps, "voided_slots": res_pl["cap_redos"],
            "stage_ms_per_iteration": {k: v[0] / max(1, args.steps) for k, v in st_pl.items()},
            "translation_error_vs_ground_truth_m": float(np.linalg.norm(T_full[:3, 3].astype(np.float64) - T_gt[:3, 3])),
        }
        if not args.no_e2e:
            _, _, dt_first = e2e_once(pm.PointToPlaneErrorMinimizer, True)  # warm-up, like the headline e2e
            _, n_it, dt = e2e_once(pm.PointToPlaneErrorMinimizer, True)
            extra["point_to_plane"]["e2e_iterations_per_s_incl_normals_knn20"] = n_it / dt
            extra["point_to_plane"]["e2e_seconds_per_registration"] = dt
            extra["point_to_plane"]["e2e_first_call_seconds"] = dt_first

    # ---- cpu_baseline on this box's host cores (rank 0, N = 1 only) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import binding as orc
        orc.build()
        threads = host_threads()
        n_it = max(1, args.cpu_sample_iters)
        orc.icp(rd, rf, filters=[(orc.FILTER_TRIMMEDDIST, RATIO)], minimizer=orc.MIN_P2POINT, max_iterations=n_it, nthreads=threads)
        tm = orc.last_timings()
        cpu = {"value": tm["iterations"] / tm["loop_s"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d iterations of the same %d x %d workload (oracle port, OpenMP over queries; kd-tree build %.2f s excluded)"
                         % (tm["iterations"], args.points, args.points, tm["build_s"])}
        orc.icp(rd, rf, filters=[(orc.FILTER_TRIMMEDDIST, RATIO)], minimizer=orc.MIN_P2POINT, max_iterations=1, nthreads=1)
        t1 = orc.last_timings()
        cpu["single_thread_value"] = t1["iterations"] / t1["loop_s"]

    ctx.close()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu, "extra": extra, "iterations_executed": res["iterations"],
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
