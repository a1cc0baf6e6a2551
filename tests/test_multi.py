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
@pytest.mark.parametrize("grid,nb", [(24, 64), (30, 128)])
def test_distributed_factor_peer_memory(grid, nb):
    """one process per GPU (needs >= 2 GPUs): peer-mapped arenas, distributed upper tree"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = {"SPLLT_DIST_GRID": str(grid), "SPLLT_DIST_NB": str(nb)}
    r = run_ranks(2, env, 29531 + nb)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "dist_check gpu ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world,mk,nb", [(2, ("poisson3d", 20), 64), (4, ("poisson3d", 24), 64), (8, ("poisson3d", 28), 32),
                                         (3, ("elasticity3d", 8), 96), (4, ("poisson3d", 40), 256),
                                         (8, ("poisson3d", 6), 8), (2, ("poisson3d", 3), 4)])   # ranks without work / no upper tree
def test_distributed_factor_emulated_on_one_gpu(world, mk, nb):
    """The multi-GPU factorization with `world` ranks EMULATED on one GPU (one process, one stream, the
    same per-rank programs / kernels / peer addressing, enqueued in an order in which no kernel waits
    for a later one): every rank's part of the factor against a single-GPU factorization, and the
    single-GPU factor against the oracle.  Runs on the driver's single-GPU test box."""
    import ctypes as C
    import numpy as np
    import torch
    import spllt_b200 as sp
    from spllt_b200 import matrices as M
    mat = getattr(M, mk[0])(mk[1])
    n, ptr, row, val = mat
    ref = sp.SpLLT(nb=nb, ncpu=world)
    assert ref.analyse(n, ptr, row) == 0
    d_val = torch.tensor(val, device="cuda")
    ref.factor_dev(d_val.data_ptr())
    ref.wait()
    assert ref.pivot_flag() == 0
    engines = []
    for r in range(world):
        s = sp.SpLLT(nb=nb, ncpu=world)
        assert s.analyse(n, ptr, row) == 0
        s.L.spllt_b200_partition(s.akeep, s.fkeep, r, world)
        engines.append(s)
    L = ref.L
    arr = (C.c_void_p * world)(*[e.fkeep for e in engines])
    for rep in range(2):            # twice: epochs / counters must carry over
        assert L.spllt_b200_emulate_ranks_factor(arr, world, C.c_void_p(d_val.data_ptr())) == 0
    out = np.zeros(2)
    for r, e in enumerate(engines):
        assert e.pivot_flag() == 0
        assert L.spllt_b200_compare_factor(e.akeep, e.fkeep, ref.akeep, ref.fkeep, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        owners = [L.spllt_b200_node_owner(e.akeep, k + 1) for k in range(e.nnodes)]
        if any(o == r or o < 0 for o in owners):
            assert out[1] > 0 and out[0] <= 1e-12 * out[1], (r, out)
        else:                                   # a rank without any node (fewer subtrees than ranks)
            assert out[0] == 0 and out[1] == 0
    # (the multi-rank solve needs NCCL all-reduces between its phases: it runs in dist_check.py on >= 2 GPUs)


# ------------------------------------------------------------------------------------------
# Host-side replay of the multi-GPU factorization programs (no GPU): every rank's launch list is
# walked in order and checked against the dependency rules of the distributed upper tree.
def _rank_tables(mat, nb, rank, world):
    import ctypes as C
    import numpy as np
    import spllt_b200 as sp
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=world)
    assert s.analyse(n, ptr, row) == 0
    L = s.L
    L.spllt_b200_partition_host(s.akeep, rank, world)
    llp = C.POINTER(C.c_longlong)
    nrec = L.spllt_b200_num_launch_records(s.akeep)
    rec = np.zeros((max(nrec, 1), 8), np.int64)
    L.spllt_b200_get_launch_records(s.akeep, rec.ctypes.data_as(llp))
    nt = L.spllt_b200_num_tile_tasks(s.akeep)
    tt = np.zeros((max(nt, 1), 10), np.int64)
    L.spllt_b200_get_tile_tasks(s.akeep, tt.ctypes.data_as(llp))
    ns = L.spllt_b200_num_top_steps(s.akeep)
    steps = np.zeros((max(ns, 1), 4), np.int32)
    L.spllt_b200_get_top_steps(s.akeep, steps.ctypes.data_as(C.POINTER(C.c_int)))
    own = np.zeros(max(s.nbcol, 1), np.int32)
    L.spllt_b200_get_bcol_owner(s.akeep, own.ctypes.data_as(C.POINTER(C.c_int)))
    return s, rec[:nrec], tt[:nt], steps[:ns], own[:s.nbcol], float(L.spllt_b200_tile_flops_algo(s.akeep))


@pytest.mark.parametrize("world,grid,nb", [(2, 14, 32), (3, 16, 48), (4, 18, 64), (8, 20, 32), (8, 6, 8), (5, 7, 16), (2, 3, 4)])
def test_distributed_programs_replay(world, grid, nb):
    import numpy as np
    from spllt_b200 import matrices as M
    mat = M.poisson3d(grid)
    ranks = [_rank_tables(mat, nb, r, world) for r in range(world)]
    s0, _, _, steps, own, _ = ranks[0]
    one = _rank_tables(mat, nb, 0, 1)
    nodes = s0.nodes()
    sptr, sparent, rptr, rlist = s0.symbolic()
    nn = s0.nnodes
    node_owner = np.array([s0.L.spllt_b200_node_owner(s0.akeep, k + 1) for k in range(nn)])
    ncols = (nodes[:, 1] - nodes[:, 0] + 1).astype(np.int64)
    nbc = -(-ncols // nb)
    bcol0 = np.concatenate([[0], np.cumsum(nbc)])[:-1]
    col2node = np.repeat(np.arange(nn), ncols)
    # the steps and the ownership table are identical on every rank; owners are dealt cyclically
    for r in range(1, world):
        assert np.array_equal(ranks[r][3], steps) and np.array_equal(ranks[r][4], own)
    assert len(steps) > 0 or grid <= 3        # (a single supernode: one rank owns everything, no upper tree)
    step_of = {}
    for t, (node, c, o, slot) in enumerate(steps):
        assert node_owner[node] == -1 and o == t % world and own[bcol0[node] + c] == o
        step_of[int(bcol0[node] + c)] = t
    assert len(step_of) == int(nbc[node_owner < 0].sum())          # every upper-tree block column has a step
    # a block column's step follows every step of its descendants in the upper tree and its own predecessors
    for t, (node, c, o, slot) in enumerate(steps):
        if c > 0:
            assert step_of[int(bcol0[node] + c - 1)] < t
        p = int(sparent[node]) - 1
        if p < nn:
            assert step_of[int(bcol0[p])] > step_of[int(bcol0[node] + nbc[node] - 1)]

    def dest_bcol(node, j0, src):
        if src < 0:
            return int(bcol0[node] + j0 // nb)
        piv = int(rlist[rptr[node] - 1 + j0]) - 1              # pivot column (0-based) of source row j0
        a = int(col2node[piv])
        return int(bcol0[a] + (piv - (sptr[a] - 1)) // nb)

    total_algo = 0.0
    pushed = set()
    for r, (s, rec, tt, _, _, algo) in enumerate(ranks):
        total_algo += algo
        avail, started = set(), set()
        cur_own = None
        last_depth = -1
        for kind, depth, begin, count, phase, tag, stream, deadline in rec:
            if phase == 0:
                # phase 0 only touches this rank's subtrees as sources
                if kind in (1, 2):
                    for t in tt[begin:begin + count]:
                        assert node_owner[t[0]] == r
                continue
            assert depth >= last_depth                             # steps in order
            last_depth = depth
            node, c, o, slot = steps[depth] if depth < len(steps) else (None, None, None, None)
            if kind == 0:                                          # panel of the step's block column
                assert o == r
                g = int(bcol0[node] + c)
                started.add(g)
                cur_own = g
            elif kind == 3:                                        # push
                assert o == r and int(begin) == int(bcol0[node] + c)
                avail.add(int(begin))
                pushed.add(int(begin))
            elif kind == 4:                                        # wait
                assert o != r and int(begin) == int(bcol0[node] + c)
                avail.add(int(begin))
            else:
                for t in tt[begin:begin + count]:
                    tnode, i0, j0, k0, mt, nt, kk, src = (int(x) for x in t[:8])
                    assert node_owner[tnode] == -1
                    d = dest_bcol(tnode, j0, src)
                    assert own[d] == r, "owner computes"
                    # destination columns stay inside one block column
                    assert dest_bcol(tnode, j0 + nt - 1, src) == d
                    if src < 0:
                        sb = int(bcol0[tnode] + k0 // nb)
                        assert (k0 + kk - 1) // nb == k0 // nb
                        if sb == d:
                            assert d == cur_own and tag in (3, 4)  # inner update of the running chain
                            continue
                        srcs = [sb]
                    else:
                        assert k0 == 0 and kk == ncols[tnode]
                        srcs = [int(bcol0[tnode] + q) for q in range(int(nbc[tnode]))]
                    for sb in srcs:
                        assert sb in avail, "source block column used before it was factorized / received"
                    assert d not in started, "update into a block column whose panel chain has started"
        assert len(started) == sum(1 for t in steps if t[2] == r)
    assert pushed == set(step_of)                                   # every upper-tree block column is pushed once
    # nothing is lost or done twice: the algorithmic flops of all ranks' tile updates add up to one GPU's
    assert abs(total_algo - one[5]) <= 1e-9 * one[5]
