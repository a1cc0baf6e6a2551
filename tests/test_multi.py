"""N > 1 path: world_size-2 gloo run on CPU (host logic: subtree -> rank mapping, per-rank work
lists) and, on a box with >= 2 GPUs, the real NCCL run checked against a single-GPU factor."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ranks(nproc, env_extra, port):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    return subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)


@pytest.mark.parametrize("world,grid,nb", [(2, 14, 32), (4, 18, 64)])
def test_partition_gloo_cpu(world, grid, nb):
    r = run_ranks(world, {"SPLLT_DIST_CPU": "1", "SPLLT_DIST_GRID": str(grid), "SPLLT_DIST_NB": str(nb),
                          "CUDA_VISIBLE_DEVICES": ""}, 29511 + world)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "dist_check cpu ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("replicated", ["0", "1"], ids=["distributed-top", "replicated-top"])
def test_distributed_factor_nccl(replicated):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = {"SPLLT_DIST_GRID": "24", "SPLLT_DIST_NB": "64"}
    if replicated == "1":
        env["SPLLT_B200_REPLICATED_TOP"] = "1"
    r = run_ranks(2, env, 29531 + int(replicated))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "dist_check gpu ok" in r.stdout
