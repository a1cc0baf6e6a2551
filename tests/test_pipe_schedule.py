"""Pipelined solve (spllt_b200/csrc/solve_pipe.cu): the value-independent work lists.

The persistent kernels let CTAs claim tasks in list order and wait on flags / counters, so the
lists must be topological: replaying them ONE task at a time must find every wait already
satisfied (=> in-order claiming cannot deadlock), every counter must reach exactly its expected
value before it is consumed, and the tasks must cover every row of every node exactly once.
A second test replays the same lists numerically (numpy, the arithmetic each task performs, on
the oracle's factor) and checks the solution -- the decomposition itself, independent of CUDA.
CPU only."""
import numpy as np
import pytest

import spllt_b200 as sp
from oracle.oracle import Oracle, chkerr
from spllt_b200 import matrices as M
from tests.cases import SMALL, MEDIUM, ids

PS = 64
DIAG, BELOW, SMALLK = 0, 1, 2

EXTRA = [
    ("p3d-24-nb64", lambda: M.poisson3d(24), 64, 4, 1),
    ("el3d-9-nb96", lambda: M.elasticity3d(9), 96, 4, 1),
    ("rand300-dense-nb32", lambda: M.random_spd(300, 0.3, 2), 32, 2, 1),
]
CASES = SMALL + MEDIUM + EXTRA


def tables(case):
    name, mk, nb, ncpu, prune = case
    n, ptr, row, val = mk()
    s = sp.SpLLT(nb=nb, ncpu=ncpu, prune_tree=prune)
    assert s.analyse(n, ptr, row) == 0
    sptr, sparent, rptr, rlist = s.symbolic()
    return s, (n, ptr, row, val), (sptr, sparent, rptr, rlist)


@pytest.mark.parametrize("case", CASES, ids=ids(CASES))
def test_lists_are_topological_and_cover(case):
    s, (n, ptr, row, val), (sptr, sparent, rptr, rlist) = tables(case)
    tf, tb, nd, dest, nstrips, expect = s.pipe_tables()
    nn = s.nnodes
    m_, n_, sa, strip0, np_, exp_f, exp_b, pflag = nd.T
    assert np.array_equal(n_, np.diff(sptr)) and np.array_equal(m_, np.diff(rptr))
    assert np.array_equal(sa, sptr[:-1] - 1)
    assert np.array_equal(np_, (n_ + PS - 1) // PS)
    assert np.array_equal(strip0, np.concatenate([[0], np.cumsum(np_)[:-1]])) and nstrips == np_.sum()
    col2node = np.repeat(np.arange(nn), n_)
    col2strip = strip0[col2node] + (np.arange(n) - sa[col2node]) // PS     # pivot column -> global strip
    idx = [rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1 for k in range(nn)]

    def strips_of(k, r0, r1):
        st = col2strip[idx[k][r0:r1]]
        return st[np.concatenate([[True], np.diff(st) != 0])] if len(st) else st

    # ---- forward: a strip reads its right-hand side only after every task that adds into its
    # rows has finished (counter == expect at that moment), and waits only on earlier tasks
    flags = np.zeros(nstrips, bool)
    cnt = np.zeros(nstrips, np.int64)
    rows_done = [np.zeros(m_[k], np.int32) for k in range(nn)]

    def fwd_below(k, r0, r1, db, dc):
        assert flags[strip0[k]:strip0[k] + np_[k]].all()          # x of the node is complete
        want = strips_of(k, r0, r1)
        assert np.array_equal(dest[db:db + dc], want)
        assert not flags[want].any()                              # lands before the ancestor strip reads
        cnt[want] += 1
        rows_done[k][r0:r1] += 1

    for node, kind, r0, nrows, db, dc in tf:
        if kind in (DIAG, SMALLK):
            i = r0 if kind == DIAG else 0
            f = strip0[node] + i
            assert flags[strip0[node]:f].all() and not flags[f]
            assert cnt[f] == expect[f]
            flags[f] = True
            rows_done[node][i * PS:min((i + 1) * PS, n_[node])] += 1
            if kind == SMALLK:
                assert np_[node] == 1
                fwd_below(node, n_[node], m_[node], db, dc)
        else:
            assert r0 >= n_[node] and 0 < nrows <= (512 if np_[node] <= 4 else PS)
            fwd_below(node, r0, r0 + nrows, db, dc)
    assert flags.all() and np.array_equal(cnt, expect)
    assert all((r == 1).all() for r in rows_done)

    # ---- backward: a below task gathers x only from strips that are already published
    flags[:] = False
    cntb = np.zeros(nn, np.int64)
    rows_done = [np.zeros(m_[k], np.int32) for k in range(nn)]

    def bwd_below(k, r0, r1, db, dc):
        want = strips_of(k, r0, r1)
        assert np.array_equal(dest[db:db + dc], want)
        assert flags[want].all()
        assert not flags[strip0[k]:strip0[k] + np_[k]].any()
        rows_done[k][r0:r1] += 1

    for node, kind, r0, nrows, db, dc in tb:
        if kind == BELOW:
            bwd_below(node, r0, r0 + nrows, db, dc)
            cntb[node] += 1
            continue
        if kind == SMALLK:
            bwd_below(node, n_[node], m_[node], db, dc)
            i = 0
        else:
            i = r0
            if i == np_[node] - 1:
                assert cntb[node] == exp_b[node]
        f0 = strip0[node]
        assert flags[f0 + i + 1:f0 + np_[node]].all() and not flags[f0 + i]
        flags[f0 + i] = True
        rows_done[node][i * PS:min((i + 1) * PS, n_[node])] += 1
    assert flags.all() and np.array_equal(cntb, exp_b)
    assert all((r == 1).all() for r in rows_done)


NUMERIC = [c for c in CASES if c[0] in ("tri3", "n1", "p2d-20-nb16", "p3d-10-nb32", "p3d-9x7x5-nb33",
                                        "rand300-dense", "el3d-6-nb128", "p3d-24-nb64", "rand300-dense-nb32")]


@pytest.mark.parametrize("case", NUMERIC, ids=ids(NUMERIC))
def test_lists_replayed_numerically(case):
    """Execute the task lists with numpy on the oracle's factor: same decomposition as the CUDA
    kernels (strip gathers, below-row scatters with counters, transposed gathers)."""
    s, (n, ptr, row, val), (sptr, sparent, rptr, rlist) = tables(case)
    name, mk, nb, ncpu, prune = case
    tf, tb, nd, dest, nstrips, expect = s.pipe_tables()
    nn = s.nnodes
    m_, n_, sa, strip0, np_, exp_f, exp_b, pflag = nd.T
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu, prune=prune)
    o.factor(val, 1)
    fo = o.factor_entries()
    # node matrices (m x n) from the reference layout (block columns, rows [c nb, m) x width)
    Ls, pos = [], 0
    for k in range(nn):
        L = np.zeros((m_[k], n_[k]))
        for c0 in range(0, n_[k], nb):
            w = min(nb, n_[k] - c0)
            h = m_[k] - c0
            L[c0:, c0:c0 + w] = fo[pos:pos + h * w].reshape(h, w)
            pos += h * w
        L[:n_[k], :n_[k]] = np.tril(L[:n_[k], :n_[k]])
        Ls.append(L)
    idx = [rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1 for k in range(nn)]
    nrhs = 3
    xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)) + 0.25 * np.sin(np.arange(n))[:, None])
    b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
    porder = np.argsort(s.order[:n] - 1)          # pivot position -> variable
    xw = b[porder, :].copy()

    def strip(k, i):
        return slice(i * PS, min((i + 1) * PS, n_[k]))

    def fwd_below(k, r0, r1):
        xw[idx[k][r0:r1]] -= Ls[k][r0:r1, :] @ xw[sa[k]:sa[k] + n_[k]]

    for node, kind, r0, nrows, db, dc in tf:
        if kind != BELOW:
            sl = strip(node, r0 if kind == DIAG else 0)
            L = Ls[node]
            rhs = xw[sa[node] + sl.start:sa[node] + sl.stop] - L[sl, :sl.start] @ xw[sa[node]:sa[node] + sl.start]
            xw[sa[node] + sl.start:sa[node] + sl.stop] = np.linalg.solve(L[sl, sl], rhs)
            if kind == SMALLK:
                fwd_below(node, n_[node], m_[node])
        else:
            fwd_below(node, r0, r0 + nrows)

    def bwd_below(k, r0, r1):
        xw[sa[k]:sa[k] + n_[k]] -= Ls[k][r0:r1, :].T @ xw[idx[k][r0:r1]]

    for node, kind, r0, nrows, db, dc in tb:
        if kind == BELOW:
            bwd_below(node, r0, r0 + nrows)
            continue
        if kind == SMALLK:
            bwd_below(node, n_[node], m_[node])
        sl = strip(node, r0 if kind == DIAG else 0)
        L = Ls[node]
        lo, hi = sa[node] + sl.stop, sa[node] + n_[node]
        rhs = xw[sa[node] + sl.start:sa[node] + sl.stop] - L[sl.stop:n_[node], sl].T @ xw[lo:hi]
        xw[sa[node] + sl.start:sa[node] + sl.stop] = np.linalg.solve(L[sl, sl].T, rhs)
    x = np.empty_like(xw)
    x[porder, :] = xw
    x = np.asfortranarray(x)
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == nrhs and err.max() <= 1e-14, err


def test_path_selection_follows_the_structure():
    """Few right-hand sides go to the persistent kernels unless most of L sits in wide nodes."""
    narrow = sp.SpLLT(nb=128)
    n, ptr, row, val = M.poisson3d(24)
    narrow.analyse(n, ptr, row)
    assert narrow.L.spllt_b200_wide_frac(narrow.akeep) < 0.6 and narrow.L.spllt_b200_pipe_max_nrhs(narrow.akeep) == 8
    wide = sp.SpLLT(nb=64)
    n = 400                                            # a dense matrix: one supernode, 7 strips wide
    ptr = np.concatenate([[1], 1 + np.cumsum(np.arange(n, 0, -1))]).astype(np.int32)
    row = np.concatenate([np.arange(j + 1, n + 1) for j in range(n)]).astype(np.int32)
    wide.analyse(n, ptr, row)
    assert wide.L.spllt_b200_wide_frac(wide.akeep) > 0.6 and wide.L.spllt_b200_pipe_max_nrhs(wide.akeep) == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_lists(world):
    """Multi-GPU solve (spllt_b200/dist.py): for every rank, the lists of its own subtrees and of
    the shared upper tree replay without an unsatisfied wait -- forward: subtrees, (all-reduce),
    upper tree with fresh counters; backward: upper tree, then the subtrees on the SAME flags --
    and over all ranks every row of every node is swept exactly once (upper tree: once per rank)."""
    n, ptr, row, val = M.poisson3d(22)
    covered = None
    for rank in range(world):
        s = sp.SpLLT(nb=64, ncpu=world)
        assert s.analyse(n, ptr, row) == 0
        s.L.spllt_b200_partition_host(s.akeep, rank, world)
        nn = s.nnodes
        own = np.array([s.L.spllt_b200_node_owner(s.akeep, k + 1) for k in range(nn)])
        tf, tb, nd, dest, nstrips, exp_own = s.pipe_tables()
        tft, tbt, exp_top = s.pipe_top_tables()
        m_, n_, sa, strip0, np_, exp_f, exp_b, pflag = nd.T
        strip_node = np.repeat(np.arange(nn), np_)
        assert set(tf[:, 0]) <= set(np.nonzero(own == rank)[0]) and set(tft[:, 0]) <= set(np.nonzero(own < 0)[0])
        rows_done = [np.zeros(m_[k], np.int32) for k in range(nn)]

        def forward(tasks, expect, in_list):
            flags = np.zeros(nstrips, bool)
            cnt = np.zeros(nstrips, np.int64)
            for node, kind, r0, nrows, db, dc in tasks:
                assert in_list[node]
                if kind != BELOW:
                    f = strip0[node] + (r0 if kind == DIAG else 0)
                    assert flags[strip0[node]:f].all() and not flags[f] and cnt[f] == expect[f]
                    flags[f] = True
                    i = f - strip0[node]
                    rows_done[node][i * PS:min((i + 1) * PS, n_[node])] += 1
                if kind != DIAG:
                    assert flags[strip0[node]:strip0[node] + np_[node]].all()
                    d = dest[db:db + dc]
                    d = d[in_list[strip_node[d]]]          # strips of other lists: summed by the all-reduce
                    assert not flags[d].any()
                    cnt[d] += 1
                    lo, hi = (r0, r0 + nrows) if kind == BELOW else (n_[node], m_[node])
                    rows_done[node][lo:hi] += 1
            assert np.array_equal(cnt[in_list[strip_node]], expect[in_list[strip_node]])
            return flags

        def backward(tasks, in_list, flags):
            cntb = np.zeros(nn, np.int64)
            for node, kind, r0, nrows, db, dc in tasks:
                assert in_list[node]
                if kind != DIAG:
                    assert flags[dest[db:db + dc]].all()        # incl. upper-tree strips raised by the previous launch
                    cntb[node] += kind == BELOW
                if kind != BELOW:
                    i = r0 if kind == DIAG else 0
                    f0 = strip0[node]
                    if kind == DIAG and i == np_[node] - 1:
                        assert cntb[node] == exp_b[node]
                    assert flags[f0 + i + 1:f0 + np_[node]].all() and not flags[f0 + i]
                    flags[f0 + i] = True
            return flags

        forward(tf, exp_own, own == rank)
        forward(tft, exp_top, own < 0)
        flags = backward(tbt, own < 0, np.zeros(nstrips, bool))
        flags = backward(tb, own == rank, flags)
        mine = (own == rank) | (own < 0)
        assert flags[mine[strip_node]].all() and not flags[~mine[strip_node]].any()
        for k in range(nn):
            assert (rows_done[k] == (1 if mine[k] else 0)).all()
        c = np.where(own == rank, 1, 0)
        covered = c if covered is None else covered + c
        if rank == 0:
            top = own < 0
    assert np.array_equal(covered + top, np.ones(nn, int))


@pytest.mark.parametrize("world", [2, 3])
def test_multi_gpu_lists_replayed_numerically(world):
    """The multi-GPU solve in numpy: every rank executes its own lists on its own copy of the work
    vector, the copies are summed where the real code all-reduces, and the result solves A x = b."""
    nb = 48
    n, ptr, row, val = M.poisson3d(14)
    ranks = []
    for rank in range(world):
        s = sp.SpLLT(nb=nb, ncpu=world)
        assert s.analyse(n, ptr, row) == 0
        s.L.spllt_b200_partition_host(s.akeep, rank, world)
        ranks.append(s)
    s0 = ranks[0]
    sptr, sparent, rptr, rlist = s0.symbolic()
    nn = s0.nnodes
    o = Oracle(n, ptr, row, s0.order, sptr, sparent, rptr, rlist, nb, ncpu=world)
    o.factor(val, 1)
    fo = o.factor_entries()
    nd = s0.pipe_tables()[2]
    m_, n_, sa = nd[:, 0], nd[:, 1], nd[:, 2]
    Ls, pos = [], 0
    for k in range(nn):
        L = np.zeros((m_[k], n_[k]))
        for c0 in range(0, n_[k], nb):
            w, h = min(nb, n_[k] - c0), m_[k] - c0
            L[c0:, c0:c0 + w] = fo[pos:pos + h * w].reshape(h, w)
            pos += h * w
        L[:n_[k], :n_[k]] = np.tril(L[:n_[k], :n_[k]])
        Ls.append(L)
    idx = [rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1 for k in range(nn)]
    col2node = np.repeat(np.arange(nn), n_)
    xs = np.asfortranarray(1.0 + 0.5 * np.sin(np.arange(n)))[:, None]
    b = np.asfortranarray(M.matvec(n, ptr, row, val, np.asfortranarray(xs)))
    porder = np.argsort(s0.order[:n] - 1)

    def fwd(xw, tasks):
        for node, kind, r0, nrows, db, dc in tasks:
            L, a = Ls[node], sa[node]
            if kind != BELOW:
                i = r0 if kind == DIAG else 0
                sl = slice(i * PS, min((i + 1) * PS, n_[node]))
                rhs = xw[a + sl.start:a + sl.stop] - L[sl, :sl.start] @ xw[a:a + sl.start]
                xw[a + sl.start:a + sl.stop] = np.linalg.solve(L[sl, sl], rhs)
            if kind != DIAG:
                lo, hi = (r0, r0 + nrows) if kind == BELOW else (n_[node], m_[node])
                xw[idx[node][lo:hi]] -= L[lo:hi, :] @ xw[a:a + n_[node]]

    def bwd(xw, tasks):
        for node, kind, r0, nrows, db, dc in tasks:
            L, a = Ls[node], sa[node]
            if kind != DIAG:
                lo, hi = (r0, r0 + nrows) if kind == BELOW else (n_[node], m_[node])
                xw[a:a + n_[node]] -= L[lo:hi, :].T @ xw[idx[node][lo:hi]]
            if kind != BELOW:
                i = r0 if kind == DIAG else 0
                sl = slice(i * PS, min((i + 1) * PS, n_[node]))
                rhs = xw[a + sl.start:a + sl.stop] - L[sl.stop:n_[node], sl].T @ xw[a + sl.stop:a + n_[node]]
                xw[a + sl.start:a + sl.stop] = np.linalg.solve(L[sl, sl].T, rhs)

    keep, copies = [], []
    for rank, s in enumerate(ranks):
        own = np.array([s.L.spllt_b200_node_owner(s.akeep, k + 1) for k in range(nn)])
        k_ = (own[col2node] == rank) | ((own[col2node] < 0) & (rank == 0))
        keep.append(k_)
        xw = b[porder, :].copy()
        xw[~k_] = 0.0                                        # phase 0
        fwd(xw, s.pipe_tables()[0])                          # phase 1
        copies.append(xw)
    total = sum(copies)                                       # all-reduce
    copies = []
    for rank, s in enumerate(ranks):
        xw = total.copy()
        tft, tbt, _ = s.pipe_top_tables()
        fwd(xw, tft)                                          # phase 2
        bwd(xw, tbt)                                          # phase 3
        bwd(xw, s.pipe_tables()[1])                           # phase 4
        xw[~keep[rank]] = 0.0
        copies.append(xw)
    xw = sum(copies)                                          # all-reduce
    x = np.empty_like(xw)
    x[porder, :] = xw
    ok, err = chkerr(n, ptr, row, val, np.asfortranarray(x), b)
    assert ok == 1 and err.max() <= 1e-14, err
    assert np.abs(x - xs).max() <= 1e-10


ORDERS = {
    "depth-node": {"SPLLT_B200_PIPE_ORDER_EST": "0", "SPLLT_B200_PIPE_BWD_EARLY": "0"},
    "backward-readiness": {"SPLLT_B200_PIPE_ORDER_EST": "0", "SPLLT_B200_PIPE_BWD_EARLY": "1"},
    "earliest-start": {"SPLLT_B200_PIPE_ORDER_EST": "1"},
}


@pytest.mark.parametrize("order", sorted(ORDERS))
@pytest.mark.parametrize("case", [SMALL[8], SMALL[11], MEDIUM[1], EXTRA[1]], ids=ids([SMALL[8], SMALL[11], MEDIUM[1], EXTRA[1]]))
def test_every_list_order_is_topological_and_exact(case, order, monkeypatch):
    """The three orders of the task lists -- (depth, node), backward BELOW tasks behind the last strip
    they read, and the earliest-start order that is the default on one GPU -- all pass the replay (no
    task waits on a later one, counters exact, rows covered once) and give the same solution."""
    for k, v in ORDERS[order].items():
        monkeypatch.setenv(k, v)
    test_lists_are_topological_and_cover(case)
    test_lists_replayed_numerically(case)


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_lists_in_earliest_start_order(world, monkeypatch):
    """the per-rank lists (own subtrees / upper tree) sorted the same way stay valid (not the default yet)"""
    monkeypatch.setenv("SPLLT_B200_PIPE_ORDER_EST", "1")
    test_multi_gpu_lists(world)
    if world == 2:
        test_multi_gpu_lists_replayed_numerically(world)


def test_full_size_lists_replay_p3d64():
    """BASELINE configs[1] (Poisson 64^3, nb = 512): the earliest-start ordered lists of the size the GPU
    runs (20 000+ tasks) pass the same replay -- no task ever waits on a later one."""
    test_lists_are_topological_and_cover(("p3d-64-nb512", lambda: M.poisson3d(64), 512, 1, 1))
