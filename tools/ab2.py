"""A/B of tuning builds on one GPU box: for every variant (libpmgpu_<name>.so, see build.py; "base" = libpmgpu.so) the
stage times of the c2plane and c4 loops, the resident normals and the staged kNN throughput.

    python tools/ab2.py base narrow onephase          (parent: one subprocess per variant)
    python tools/ab2.py base env:PMGPU_NO_FUSED_SELECT=1   (run-time switches of the base build)
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(configs):
    import numpy as np
    import torch
    import bench
    from libpointmatcher_b200 import capi, pm
    from libpointmatcher_b200 import dist as pmdist
    torch.cuda.set_device(0)
    out = {}
    for name in configs:
        args = argparse.Namespace(gpus=1, steps=20, warmup=5, impl="ours", config=name, points=0, pairs=0, mode="auto", reps=5, e2e_reps=3, rings=0,
                                  cpu_sample_iters=1, no_extra=True, no_cpu=True, no_e2e=True)
        cfg = bench.resolved(args)
        m = bench.measure(args, cfg, torch, capi, pm, pmdist, 0, 1, 0, "single", args.reps, want_e2e=False)
        r = {"it_per_s": round(m["value"], 1), "ms_per_step": round(m["ms_per_step"], 4), "first_ms": round(m["first_iteration_ms"], 3),
             "last_ms": round(m["last_iteration_ms"], 3), "b2b_ms": round(m["back_to_back_ms_per_step"], 4),
             "stage": {k: round(v, 4) for k, v in m["stage_ms_per_iteration"].items()}, "normals_ms": round(m["normals_ms"] or 0, 3)}
        # fingerprint of the RESULTS: variants must agree on it, or the faster one is computing something else
        import hashlib
        T_bits = m["tm"].ctx.icp_result()["T_iter"].tobytes()
        ids, dists, _ = m["ctx"].knn(m["tm"].ctx.icp_result()["T_iter"], cfg["knn"], 0.0, cfg["max_dist"])
        r["fingerprint"] = hashlib.sha1(T_bits + dists.tobytes()).hexdigest()[:12]
        if name == "c2plane":
            r["knn"] = {k: {a: round(b, 3) for a, b in v.items() if a.endswith("_ms")} for k, v in bench.knn_throughput(args, m, capi, 0).items() if k.startswith("k")}
        m["ctx"].close()
        out[name] = r
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2].split(","))
        sys.exit(0)
    configs = os.environ.get("AB_CONFIGS", "c2plane,c4")
    for v in sys.argv[1:]:
        env = dict(os.environ)
        if v == "base":
            env.pop("PMGPU_VARIANT", None)
        elif v.startswith("env:"):   # env:NAME=value[,NAME=value]: the base build with run-time switches
            env.pop("PMGPU_VARIANT", None)
            for kv in v[4:].split(","):
                name, _, val = kv.partition("=")
                env[name] = val
        else:
            env["PMGPU_VARIANT"] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", configs], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        print("%-10s rc=%d %s" % (v, r.returncode, line if r.returncode == 0 else r.stderr[-1500:]), flush=True)
