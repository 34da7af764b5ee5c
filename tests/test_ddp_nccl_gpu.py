"""2-GPU NCCL run of the CUDA training step against shard-and-average (see tests/ddp_nccl_worker.py).
Needs two visible GPUs (`gpurun --gpus 2`); skipped on a single-GPU box.  The log is kept in gpurun_out/."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ddp_step_matches_shard_and_average(cuda_dev):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ddp_nccl_parity.log"), "w") as fh:
        fh.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    lines = [json.loads(l.split(" ", 1)[1]) for l in r.stdout.splitlines() if l.startswith("DDP_NCCL_RESULT ")]
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert len(lines) == world and all(l["ok"] for l in lines)
