"""Bit-exact parity of the value-independent structures (SURVEY.md 8a 'indexing structures'):
the product's closed-form tables vs the oracle's restatement of the reference loops, from the
same (order, sptr, sparent, rptr, rlist).  CPU only (spllt_analyse is host code)."""
import numpy as np
import pytest

import spllt_b200 as sp
from oracle.oracle import Oracle
from tests.cases import SMALL, MEDIUM, ids


def analysed(case):
    name, mk, nb, ncpu, prune = case
    n, ptr, row, val = mk()
    s = sp.SpLLT(nb=nb, ncpu=ncpu, prune_tree=prune)
    assert s.analyse(n, ptr, row) == 0
    sptr, sparent, rptr, rlist = s.symbolic()
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu, prune=prune)
    return s, o, (n, ptr, row, val)


@pytest.mark.parametrize("case", SMALL + MEDIUM, ids=ids(SMALL + MEDIUM))
def test_tables_bit_exact(case):
    s, o, _ = analysed(case)
    assert np.array_equal(s.nodes(), o.nodes())
    assert np.array_equal(s.blocks(), o.blocks())
    assert np.array_equal(s.weight(), o.weight())
    assert np.array_equal(s.small(), o.small())
    assert s.nbcol == o.nbcol()
    assert s.L.spllt_b200_maxmn(s.akeep) == o.maxmn()
    for b in range(1, s.nbcol + 1):
        d1, s1 = s.lmap(b)
        d2, s2 = o.lmap(b)
        assert np.array_equal(d1, d2) and np.array_equal(s1, s2)
    o.prepare_solve(3)
    assert np.array_equal(s.sblocks(), o.sblocks())
    assert s.prepare_solve_size(3) == o.L.orc_prepare_solve(o.h, o.nb, 3)


@pytest.mark.parametrize("case", SMALL, ids=ids(SMALL))
def test_symbolic_invariants(case):
    """The SSIDS stand-in: order is a permutation, supernodes partition the columns, row lists
    are sorted, start with the node's own columns, nest along the tree, and reproduce the
    column counts of a dense symbolic factorization."""
    s, o, (n, ptr, row, val) = analysed(case)
    sptr, sparent, rptr, rlist = s.symbolic()
    nn = s.nnodes
    assert sorted(s.order[:n]) == list(range(1, n + 1))
    assert sptr[0] == 1 and sptr[-1] == n + 1 and np.all(np.diff(sptr) > 0)
    assert np.all(sparent > np.arange(1, nn + 1))
    for k in range(nn):
        idx = rlist[rptr[k] - 1:rptr[k + 1] - 1]
        nc = sptr[k + 1] - sptr[k]
        assert np.array_equal(idx[:nc], np.arange(sptr[k], sptr[k + 1]))
        assert np.all(np.diff(idx) > 0)
        if nc < len(idx):
            p = sparent[k]
            assert p <= nn and sptr[p - 1] <= idx[nc] < sptr[p]   # parent owns the first row below
            pidx = set(rlist[rptr[p - 1] - 1:rptr[p] - 1])
            assert set(idx[nc:]) <= pidx
    if n <= 1000:
        # dense symbolic Cholesky of the permuted pattern: struct(L) must be inside the supernodal
        # structure, and identical when no amalgamation fill is allowed (checked as a subset here)
        a = np.zeros((n, n), bool)
        cols = np.repeat(np.arange(n), np.diff(ptr))
        pr, pc = s.order[row - 1] - 1, s.order[cols] - 1
        a[np.maximum(pr, pc), np.minimum(pr, pc)] = True
        for j in range(n):
            r = np.nonzero(a[j + 1:, j])[0] + j + 1
            if len(r):
                a[np.ix_(r, r)] |= np.tril(np.ones((len(r), len(r)), bool))
        col2node = np.repeat(np.arange(nn), np.diff(sptr))
        for j in range(n):
            k = col2node[j]
            idx = rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1
            assert set(np.nonzero(a[j:, j])[0] + j) <= set(idx[idx >= j])


def test_options_defaults():
    o = sp.Options()
    assert (o.nb, o.nemin, o.prune_tree, o.min_width_blas, o.ncpu, o.chunk) == (16, 32, 1, 8, 1, 10)


def test_nb_default_when_nonpositive():
    from spllt_b200 import matrices as M
    n, ptr, row, val = M.poisson2d(12)
    s = sp.SpLLT(nb=0)
    s.analyse(n, ptr, row)
    assert np.all(s.nodes()[:, 5] == 256)   # nb_default, src/spllt_data_mod.F90:39


def test_oracle_side_symbolic_front_end_matches_product():
    """bench.py's CPU arm takes order / sptr / sparent / rptr / rlist from oracle/libssids_standin.so (the
    SSIDS stand-in compiled into its own shared object so that the arm never loads the product): same
    tables as the product's spllt_analyse, bit for bit."""
    import numpy as np
    import spllt_b200 as sp
    from spllt_b200 import matrices as M
    from oracle import oracle as O
    for mk, nb in ((lambda: M.poisson3d(14), 32), (lambda: M.elasticity3d(5), 24), (lambda: M.poisson2d(40), 16)):
        n, ptr, row, val = mk()
        order, sptr, sparent, rptr, rlist = O.symbolic(n, ptr, row, nemin=32)
        s = sp.SpLLT(nb=nb)
        assert s.analyse(n, ptr, row) == 0
        a = s.symbolic()
        assert np.array_equal(order, s.order[:n])
        for x, y in zip((sptr, sparent, rptr, rlist), a):
            assert np.array_equal(x, y)
