"""Two-rank NCCL run of the sharded entry points on real GPUs (marker gpu; skipped on a one-GPU box).

The reference's multi-worker mode gives the serial result bit for bit (monte_carlo.py:126-139); the claim
checked here on hardware is the device analogue: sharded == single-device (statistics to 1e-9, CRR and
item-sharded book prices bit-identical), with the reduction done by the C-ABI collective
(optp_allreduce_stats).  The host-side sharding logic alone is covered on CPU by
tests/test_distributed_cpu.py (gloo, world_size 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_nccl_sharded_equals_single_device(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); the one-GPU box runs the gloo tests instead")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    report = json.loads(out.read_text())
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, "multigpu_parity.json"), "w") as f:
            json.dump(report, f, indent=1)
    assert report["world"] == 2 and report["all_ranks_ok"], report
    bad = [c for c in report["checks"] if not c["ok"]]
    assert not bad, bad
    assert len(report["checks"]) >= 15
