"""Multi-rank check, launched by torchrun (one rank per GPU; backend nccl) or, with
SPLLT_DIST_CPU=1, on CPU with backend gloo (host logic only).

GPU mode: every rank runs the distributed factorization (subtree mapping, generated elements
scattered into the owners' HBM over peer-mapped memory, upper tree distributed by block column)
and a plain single-GPU factorization of the same matrix, and compares the factor entries of every
block column it holds (its own subtrees + the upper tree); then the distributed solve and the
reporting of a non-positive pivot on every rank.
CPU mode: the ranks exchange their partition tables over gloo and check that they agree, that
every pruned subtree has exactly one owner and that the per-rank work lists cover every block
column exactly once.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spllt_b200 as sp  # noqa: E402
from spllt_b200 import matrices as M  # noqa: E402
from spllt_b200.dist import DistSpLLT  # noqa: E402


def main():
    cpu = os.environ.get("SPLLT_DIST_CPU", "0") == "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if cpu:
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = int(os.environ.get("SPLLT_DIST_GRID", "20"))
    nb = int(os.environ.get("SPLLT_DIST_NB", "64"))
    n, ptr, row, val = M.poisson3d(grid)
    stream = None
    if not cpu:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
    d = DistSpLLT(nb=nb, rank=rank, world=world, stream=stream)
    s = d.local
    if cpu:
        # host part of analyse + partition only (no device)
        s.analyse(n, ptr, row)
        s.L.spllt_b200_partition_host(s.akeep, rank, world)
    else:
        d.analyse(n, ptr, row)
    nn = s.nnodes
    own = np.array([s.L.spllt_b200_node_owner(s.akeep, k + 1) for k in range(nn)], dtype=np.int64)
    small = s.small().astype(np.int64)
    # ---- partition tables agree on all ranks
    t = torch.from_numpy(own.copy())
    gathered = [torch.zeros_like(t) for _ in range(world)]
    if not cpu:
        t = t.cuda()
        gathered = [g.cuda() for g in gathered]
    dist.all_gather(gathered, t)
    for g in gathered:
        assert torch.equal(g.cpu(), torch.from_numpy(own)), "ranks disagree on the subtree mapping"
    assert np.all(own < world) and np.all(own >= -1)
    par = s.nodes()[:, 2].astype(np.int64) - 1
    roots = []
    for k in range(nn):
        if own[k] >= 0:
            # an owned node is either a subtree root (parent shared / none) or has its parent's owner
            if par[k] >= nn or own[par[k]] < 0:
                roots.append(k)
                lo = int(s.nodes()[k, 4]) - 1
                assert np.all(own[lo:k + 1] == own[k]), "a subtree must have one owner"
            else:
                assert own[par[k]] == own[k]
        else:
            assert par[k] >= nn or own[par[k]] < 0, "the shared part must be the top of the tree"
    assert len(set(own[roots])) == min(world, len(roots)) or len(roots) >= world
    w = s.weight()
    loads = np.array([sum(int(w[k]) for k in roots if own[k] == r) for r in range(world)], dtype=np.float64)
    assert loads.max() <= 1.25 * max(loads.mean(), 1.0) or len(roots) <= world
    # ---- work lists: this rank's panels cover exactly its nodes + the shared ones
    cnt = np.zeros(nn, dtype=np.int64)
    s.L.spllt_b200_panel_coverage(s.akeep, cnt.ctypes.data_as(C.POINTER(C.c_longlong)))
    mine = (own == rank) | (own == -1)
    nodes = s.nodes()
    npanels = np.array([sum(-(-min(nb, int(nodes[k, 1] - nodes[k, 0] + 1) - c0) // 64)
                            for c0 in range(0, int(nodes[k, 1] - nodes[k, 0] + 1), nb)) for k in range(nn)])
    assert np.array_equal(cnt[own == rank], npanels[own == rank]), "a block column of this rank is missing / duplicated"
    assert np.all(cnt[(own >= 0) & (own != rank)] == 0), "this rank schedules work on a foreign subtree"
    # upper tree: distributed by block column -> every panel exactly once over all ranks
    tsum = torch.from_numpy(cnt.copy())
    if not cpu:
        tsum = tsum.cuda()
    dist.all_reduce(tsum)
    tsum = tsum.cpu().numpy()
    top = own == -1
    assert s.L.spllt_b200_dist_top(s.akeep) == 1
    assert np.array_equal(tsum[top], npanels[top]), "an upper-tree panel is missing / duplicated across ranks"
    if cpu:
        dist.barrier()
        if rank == 0:
            print("dist_check cpu ok: world %d, %d nodes, %d subtrees, shared flops %.0f%%" % (world, nn, len(roots), 100.0 * (1.0 - loads.sum() / float(w[-1]))))
        dist.destroy_process_group()
        return

    # ---- GPU: distributed factor vs single-GPU factor
    d_val = torch.tensor(val, device="cuda")
    d.factor_dev(d_val)
    d.wait()
    torch.cuda.synchronize()
    assert d.pivot_flag() == 0
    for rep in range(2):          # graph replay: epochs / flags / counters carry over
        d.factor_dev(d_val)
        d.wait()
    ref = sp.SpLLT(nb=nb, ncpu=world)
    ref.analyse(n, ptr, row)
    ref.factor(val)
    ref.wait()
    worst = 0.0
    blocks = s.blocks()
    for k in range(nn):
        if not mine[k]:
            continue
        ncol = int(nodes[k, 1] - nodes[k, 0] + 1)
        bcol0 = int(blocks[int(nodes[k, 6]) - 1, 7])
        for c in range(-(-ncol // nb)):
            a, b = s.lcol(bcol0 + c), ref.lcol(bcol0 + c)
            w = min(nb, ncol - c * nb)
            m = np.tril(np.ones((a.size // w, w), bool)).ravel()
            worst = max(worst, float(np.abs(a - b)[m].max() / max(np.abs(b).max(), 1e-300)))
    assert worst <= 1e-12, worst
    cmpd = d.compare_with_single_gpu(d_val)      # the device-side comparison bench.py reports
    assert cmpd["max_rel_diff"] <= 1e-12, cmpd
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # a matrix that is not positive definite is reported on every rank, whoever owns the pivot
    bad = val.copy()
    bad[ptr[n // 2] - 1] = -1.0
    d.factor_dev(torch.tensor(bad, device="cuda"))
    d.wait()
    assert d.pivot_flag() > 0
    d.factor_dev(d_val)
    d.wait()
    assert d.pivot_flag() == 0
    # ---- distributed solve (subtree sweeps on their owners, upper tree redundantly, two all-reduces
    # of the work vector) vs the single-GPU solve and the reference's backward-error gate
    serr = 0.0
    for nrhs in (1, 3):
        xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)) + 0.5 * np.cos(np.arange(n))[:, None])
        b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
        for rep in range(2):      # the second call replays with flags / counters reset
            dx = torch.tensor(b.T.copy(), device="cuda")
            d.solve_dev(dx, nrhs)
            d.wait()
            torch.cuda.synchronize()
        x = np.asfortranarray(dx.cpu().numpy().T)
        ok, err = sp.chkerr(n, ptr, row, val, x, b)
        assert ok == nrhs and err.max() <= 1e-14, err
        xr = b.copy(order="F")
        ref.prepare_solve(nrhs)
        ref.solve(xr, 0)
        assert np.max(np.abs(x - xr)) <= 1e-10 * np.abs(xr).max()
        serr = max(serr, float(err.max()))
    if rank == 0:
        print("dist_check gpu ok: world %d, max rel diff vs single-GPU factor %.2e, distributed solve bwd err %.2e; %s"
              % (world, t.item(), serr, d.describe()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
