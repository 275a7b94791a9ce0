"""SURVEY 8e rows 1-2 on real GPUs: a registration whose reading is sharded over 2 GPUs (exchanges fused into the
kernels over peer mailboxes; map normals per slice + all-gather) against the same registration on one GPU.
Needs >= 2 GPUs (`gpurun --gpus 2`); the assertions live in tests/sharded_worker.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_sharded_equals_single_gpu(exchange, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "sharded_worker.py"), "--points", "200000", "--json", str(out)]
    if exchange == "nccl":
        cmd.append("--nccl-only")
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    rep = json.load(open(out))
    assert rep["pass"] and rep["normals_bit_equal"], rep
    for name, c in rep["configs"].items():
        assert c["limits_bit_equal"] and c["iterations"][0] == c["iterations"][1] and c["rot_err"] <= 1e-5 and c["trans_err"] <= 1e-5, (name, c)
