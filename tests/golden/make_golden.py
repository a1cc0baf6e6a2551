"""Generates tests/golden/*.npz.  Run from the repo root: python tests/golden/make_golden.py

The reference cannot be built in this image (Fortran + SPRAL), so the fixtures pin
(a) the symbolic front end + tiling (order, sptr, sparent, rptr, rlist, tile table, pruning),
(b) factor entries checked against a dense LAPACK Cholesky before being written, and
(c) solutions checked against the known x* (b = A x*, x*(:, r) = r) and the 1e-14 gate.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spllt_b200 as sp  # noqa: E402
from spllt_b200 import matrices as M  # noqa: E402
from oracle.oracle import Oracle, chkerr  # noqa: E402

CASES = [
    ("tri3", M.tridiag3(), 4, 1, 1),
    ("poisson2d_12_nb8", M.poisson2d(12), 8, 2, 2),
    ("poisson3d_7_nb16", M.poisson3d(7), 16, 4, 3),
    ("elasticity3d_4_nb24", M.elasticity3d(4), 24, 2, 2),
]

for name, (n, ptr, row, val), nb, ncpu, nrhs in CASES:
    s = sp.SpLLT(nb=nb, ncpu=ncpu)
    s.analyse(n, ptr, row)
    sptr, sparent, rptr, rlist = s.symbolic()
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu)
    o.factor(val, 1)
    f = o.factor_entries()
    # (b) dense check
    a = M.to_dense(n, ptr, row, val)
    p = np.argsort(s.order[:n])
    ld = np.linalg.cholesky(a[np.ix_(p, p)])
    pos = 0
    for k in range(s.nnodes):
        sa, en = sptr[k] - 1, sptr[k + 1] - 2
        idx = rlist[rptr[k] - 1:rptr[k + 1] - 1] - 1
        for c0 in range(0, en - sa + 1, nb):
            w = min(nb, en - sa + 1 - c0)
            blk = ld[np.ix_(idx[c0:], np.arange(sa + c0, sa + c0 + w))]
            got = f[pos:pos + blk.size].reshape(blk.shape)
            pos += blk.size
            m = np.tril(np.ones(blk.shape, bool))
            assert np.abs(got - blk)[m].max() < 1e-13
    xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
    if name == "tri3":
        rhs = np.asfortranarray(np.ones((3, 1)))
    else:
        rhs = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
    x = np.asfortranarray(rhs.copy())
    o.prepare_solve(x.shape[1])
    o.solve(x, 0)
    ok, err = chkerr(n, ptr, row, val, x, rhs)
    assert ok == x.shape[1], err
    if name == "tri3":
        assert np.allclose(x[:, 0], [1.5, 2.0, 1.5], atol=4e-16)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), n=n, ptr=ptr, row=row, val=val, nb=nb,
                        ncpu=ncpu, order=s.order[:n], sptr=sptr, sparent=sparent, rptr=rptr, rlist=rlist,
                        blocks=s.blocks(), small=s.small(), factor=f, rhs=rhs, x=x)
    print(name, "n", n, "nnodes", s.nnodes, "factor entries", f.size)
