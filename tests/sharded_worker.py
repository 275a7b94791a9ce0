"""Worker of tests/test_gpu_sharded.py — one process per GPU (torchrun), SURVEY 8e rows 1-2.

Every rank runs the SHARDED registration (reading split into contiguous column ranges, reference
replicated, exchanges fused into the kernels over peer mailboxes, map normals computed per slice and
all-gathered); rank 0 also runs the same registration UNSHARDED on its GPU.  Asserted, for
  A: knn 1 / TrimmedDist 0.75 / PointToPlane
  B: knn 10, maxDist 2 / MaxDist 1 x MedianDist 3 / PointToPlaneWithCov
 - sharded map normals == unsharded map normals, bit for bit
 - staged path: the quantile limits are bit-equal, the incremental T within 1e-5 rad / 1e-5 m
 - fused loop: iteration counts equal, final T within 1e-5 rad / 1e-5 m, and every rank holds the
   same T bit for bit (all ranks sum the same contributions in the same order)
Usage: torchrun --nproc-per-node 2 tests/sharded_worker.py [--points N] [--json PATH]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=200_000)
    ap.add_argument("--json", default=None)
    ap.add_argument("--nccl-only", action="store_true", help="exchange through NCCL between kernels instead of the peer mailboxes")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from helpers import rot_angle
    from libpointmatcher_b200 import capi, synth
    from libpointmatcher_b200 import dist as pmdist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rd, rf, _ = synth.scan_pair(args.points)
    mean = rf[:, :3].mean(axis=0).astype(np.float32)
    rf = rf.copy(); rf[:, :3] -= mean
    rd = rd.copy(); rd[:, :3] -= mean
    configs = {
        "A": dict(knn=1, max_dist=np.inf, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=capi.MIN_P2PLANE, iters=12),
        "B": dict(knn=10, max_dist=2.0, filters=[(capi.FILTER_MAXDIST, 1.0), (capi.FILTER_MEDIANDIST, 3.0)], minimizer=capi.MIN_P2PLANE_COV, iters=6),
    }

    def run(ctx, reading, cfg):
        out = {}
        ctx.set_reading(reading)
        # staged
        ctx.knn(None, cfg["knn"], 0.0, cfg["max_dist"], download=False)
        _, limits = ctx.weights(cfg["filters"], download=False)
        T, cov, stats = ctx.minimize(cfg["minimizer"])
        out["limits"], out["T_staged"], out["stats_staged"] = limits.copy(), T, stats
        # fused loop
        p = capi.make_params(knn=cfg["knn"], max_dist=cfg["max_dist"], filters=cfg["filters"], minimizer=cfg["minimizer"],
                             max_iterations=cfg["iters"], differential=(1e-4, 1e-4, 3))
        res = ctx.icp_run(p)
        out["T"], out["iterations"], out["cov"], out["stats"] = res["T_iter"], res["iterations"], res["cov"], res["stats"]
        return out

    single = None
    normals_single = None
    if rank == 0:
        with capi.Context(local) as c1:
            c1.set_reference(rf)
            c1.ref_compute_normals(knn=10)
            normals_single = c1.ref_normals()
            single = {name: run(c1, rd, cfg) for name, cfg in configs.items()}
    dist.barrier()

    ctx = capi.Context(local)
    pmdist.init_comm(ctx, capi, peer=not args.nccl_only, nccl=True)
    ctx.set_reference(rf)
    ctx.ref_compute_normals(knn=10)          # this rank's slice + all-gather
    normals_sharded = ctx.ref_normals()
    sharded = {name: run(ctx, pmdist.shard_take(rd, rank, world), cfg) for name, cfg in configs.items()}

    # every rank must hold rank 0's T bit for bit
    report = {"world": world, "points": args.points, "exchange": "nccl" if args.nccl_only else "peer mailboxes", "configs": {}}
    ok = True
    for name in configs:
        t = torch.from_numpy(np.ascontiguousarray(sharded[name]["T"]).view(np.int32).copy()).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        same = all(bool((g == gathered[0]).all()) for g in gathered)
        if rank == 0:
            s1, sN = single[name], sharded[name]
            r = {
                "limits_bit_equal": bool((s1["limits"].view(np.uint32) == sN["limits"].view(np.uint32)).all()),
                "limits": [float(x) for x in sN["limits"]],
                "staged_rot_err": rot_angle(s1["T_staged"], sN["T_staged"]),
                "staged_trans_err": float(np.linalg.norm(s1["T_staged"][:3, 3].astype(np.float64) - sN["T_staged"][:3, 3])),
                "iterations": [s1["iterations"], sN["iterations"]],
                "rot_err": rot_angle(s1["T"], sN["T"]),
                "trans_err": float(np.linalg.norm(s1["T"][:3, 3].astype(np.float64) - sN["T"][:3, 3])),
                "stats_equal": s1["stats"] == sN["stats"],
                "all_ranks_same_T_bits": same,
            }
            if configs[name]["minimizer"] == capi.MIN_P2PLANE_COV:
                c1, cN = s1["cov"].astype(np.float64), sN["cov"].astype(np.float64)
                r["cov_rel_err"] = float(np.abs(c1 - cN).max() / max(np.abs(c1).max(), 1e-30))
            r["pass"] = (r["limits_bit_equal"] and r["staged_rot_err"] <= 1e-5 and r["staged_trans_err"] <= 1e-5 and
                         r["iterations"][0] == r["iterations"][1] and r["rot_err"] <= 1e-5 and r["trans_err"] <= 1e-5 and
                         r["stats_equal"] and same and r.get("cov_rel_err", 0.0) <= 2e-3)
            ok = ok and r["pass"]
            report["configs"][name] = r
    if rank == 0:
        report["normals_bit_equal"] = bool((normals_single.view(np.uint32) == normals_sharded.view(np.uint32)).all())
        ok = ok and report["normals_bit_equal"]
        report["pass"] = bool(ok)
        print(json.dumps(report), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(report, f, indent=1)
    ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
