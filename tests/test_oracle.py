"""Pins the oracle (oracle/spllt_oracle.cpp): known answers of the reference, independent dense
LAPACK results, the reference's acceptance gate, and the committed golden fixtures.  CPU only."""
import glob
import os

import numpy as np
import pytest

import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle, chkerr
from tests.cases import SMALL, MEDIUM, ids

HERE = os.path.dirname(os.path.abspath(__file__))


def setup(case, nthreads=1):
    name, mk, nb, ncpu, prune = case
    n, ptr, row, val = mk()
    s = sp.SpLLT(nb=nb, ncpu=ncpu, prune_tree=prune)
    s.analyse(n, ptr, row)
    sptr, sparent, rptr, rlist = s.symbolic()
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu, prune=prune)
    o.factor(val, nthreads)
    return s, o, (n, ptr, row, val)


def dense_factor_in_ref_layout(s, mat):
    """Dense LAPACK Cholesky of P A P^T laid out like lfact(bcol)%lcol (block columns, row-major)."""
    n, ptr, row, val = mat
    a = M.to_dense(n, ptr, row, val)
    p = np.argsort(s.order[:n])
    ld = np.linalg.cholesky(a[np.ix_(p, p)])
    sptr, sparent, rptr, rlist = s.symbolic()
    nb = s.options.nb
    out, mask = [], []
    for k in range(s.nnodes):
        sa, en = sptr[k] - 1, sptr[k + 1] - 2
        idx = rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1
        for c0 in range(0, en - sa + 1, nb):
            w = min(nb, en - sa + 1 - c0)
            blk = ld[np.ix_(idx[c0:], np.arange(sa + c0, sa + c0 + w))]
            out.append(blk.ravel())
            mask.append(np.tril(np.ones(blk.shape, bool)).ravel())   # strict upper of the diagonal tile is unused
    return np.concatenate(out), np.concatenate(mask)


def test_known_answer_simple_c():
    """example/C/simple.c:25-52: tridiag(-1, 2, -1), b = 1 -> x = [1.5, 2, 1.5]."""
    s, o, (n, ptr, row, val) = setup(SMALL[0])
    x = np.ones(3)
    o.prepare_solve(1)
    assert o.solve(x, 0) == 0
    assert np.allclose(x, [1.5, 2.0, 1.5], rtol=0, atol=4e-16)


@pytest.mark.parametrize("case", SMALL, ids=ids(SMALL))
def test_factor_matches_dense_lapack(case):
    s, o, mat = setup(case)
    ref, mask = dense_factor_in_ref_layout(s, mat)
    got = o.factor_entries()
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)[mask]) <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("case", SMALL + MEDIUM[:2], ids=ids(SMALL + MEDIUM[:2]))
@pytest.mark.parametrize("nrhs", [1, 3])
def test_solve_backward_error_gate(case, nrhs):
    """test/test_solve_phasis.F90:140-155, 245-315: rhs = A * (r * ones), err <= 1e-14; job 1 + 2 == job 0."""
    s, o, (n, ptr, row, val) = setup(case)
    xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
    b = M.matvec(n, ptr, row, val, xs)
    o.prepare_solve(nrhs)
    x = np.asfortranarray(b.copy())
    assert o.solve(x, 0) == 0
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == nrhs, err
    x2 = np.asfortranarray(b.copy())
    assert o.solve(x2, 1) == 0 and o.solve(x2, 2) == 0
    assert np.array_equal(x, x2)
    assert o.solve(x2, 6) == -10          # src/spllt_solve_mod.F90:216-220


@pytest.mark.parametrize("case", [SMALL[4], SMALL[8], MEDIUM[0]], ids=ids([SMALL[4], SMALL[8], MEDIUM[0]]))
def test_omp_task_build_matches_sequential(case):
    s, o1, mat = setup(case, 1)
    _, o4, _ = setup(case, 4)
    f1, f4 = o1.factor_entries(), o4.factor_entries()
    assert np.max(np.abs(f1 - f4)) <= 1e-13 * np.abs(f1).max()


def test_chkerr_formula():
    n, ptr, row, val = M.poisson2d(6)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    b = rng.standard_normal(n)
    a = M.to_dense(n, ptr, row, val)
    want = np.linalg.norm(b - a @ x) / (np.linalg.norm(b) + np.abs(val).max() * np.linalg.norm(x))
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == 0 and abs(err[0] - want) <= 1e-15 * want
    ok2, err2 = sp.chkerr(n, ptr, row, val, x, b)        # the product's host-side spllt_chkerr
    assert ok2 == 0 and abs(err2[0] - want) <= 1e-15 * want


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_golden_fixtures(path):
    """Committed fixtures (tests/golden/make_golden.py): symbolic tables, oracle factor and solution."""
    g = np.load(path)
    n, ptr, row, val = int(g["n"]), g["ptr"], g["row"], g["val"]
    nb, ncpu = int(g["nb"]), int(g["ncpu"])
    s = sp.SpLLT(nb=nb, ncpu=ncpu)
    s.analyse(n, ptr, row)
    sptr, sparent, rptr, rlist = s.symbolic()
    for k, v in (("order", s.order[:n]), ("sptr", sptr), ("sparent", sparent), ("rptr", rptr), ("rlist", rlist),
                 ("blocks", s.blocks()), ("small", s.small())):
        assert np.array_equal(g[k], v), k
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu)
    o.factor(val, 1)
    f = o.factor_entries()
    assert np.max(np.abs(f - g["factor"])) <= 1e-14 * np.abs(g["factor"]).max()
    x = np.asfortranarray(g["rhs"].copy())
    o.prepare_solve(x.shape[1])
    o.solve(x, 0)
    assert np.max(np.abs(x - g["x"])) <= 1e-12 * np.abs(g["x"]).max()
