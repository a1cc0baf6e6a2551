import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle, chkerr
import scipy.linalg as sla

def run(name, mat, nb, ncpu=1, nrhs=1):
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=ncpu)
    t=time.time(); s.analyse(n, ptr, row); ta=time.time()-t
    sptr, sparent, rptr, rlist = s.symbolic()
    print(name, 'n', n, 'nnodes', s.nnodes, 'nbcol', s.nbcol, 'flops %.3e'%s.num_flops, 'nfac %.3e'%s.num_factor, 'depth', s.L.spllt_b200_num_depth(s.akeep), 'analyse %.2fs'%ta)
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu)
    assert np.array_equal(s.blocks(), o.blocks()), 'blocks'
    assert np.array_equal(s.nodes(), o.nodes()), 'nodes'
    assert np.array_equal(s.weight(), o.weight()), 'weight'
    assert np.array_equal(s.small(), o.small()), ('small', s.small()[:50], o.small()[:50])
    for b in range(1, s.nbcol+1):
        d1,s1 = s.lmap(b); d2,s2 = o.lmap(b)
        assert np.array_equal(d1,d2) and np.array_equal(s1,s2), ('lmap', b)
    assert s.L.spllt_b200_maxmn(s.akeep) == o.maxmn()
    t=time.time(); o.factor(val, 1); tf=time.time()-t
    print('  oracle factor %.3fs  %.2f GF/s'%(tf, s.num_flops/tf/1e9))
    o.prepare_solve(nrhs)
    assert np.array_equal(s.sblocks(), o.sblocks()), 'sblocks'
    xs = np.asfortranarray(np.tile(np.arange(1, nrhs+1, dtype=float), (n,1)))
    b = M.matvec(n, ptr, row, val, xs)
    x = np.asfortranarray(b.copy())
    o.solve(x, 0)
    ok, err = chkerr(n, ptr, row, val, x, b)
    print('  oracle bwd err', err.max(), ok, 'fwd err', np.abs(x-xs).max())
    if n <= 3000:
        A = M.to_dense(n, ptr, row, val)
        p = np.argsort(s.order[:n])   # p[k] = variable at pivot k
        Ap = A[np.ix_(p,p)]
        Ld = np.linalg.cholesky(Ap)
        # compare factor entries
        fe = o.factor_entries(); blocks = o.blocks(); nodes=o.nodes()
        pos=0; maxerr=0
        for nd in range(s.nnodes):
            sa,en = nodes[nd,0]-1, nodes[nd,1]-1
            idx = rlist[rptr[nd]-1:rptr[nd+1]-1]-1
            for c0 in range(0, en-sa+1, nb):
                w = min(nb, en-sa+1-c0); h = len(idx)-c0
                blk = fe[pos:pos+h*w].reshape(h,w); pos+=h*w
                ref = Ld[np.ix_(idx[c0:], np.arange(sa+c0, sa+c0+w))]
                # upper triangle of diag block not meaningful
                mask = np.tril(np.ones((h,w),bool), 0)
                maxerr = max(maxerr, np.abs((blk-ref)[mask]).max())
        print('  oracle factor vs dense chol maxerr', maxerr)
    return s, o

run('tri3', M.tridiag3(), 4)
run('p2d-20', M.poisson2d(20), 16)
run('p2d-30 nb8', M.poisson2d(30), 8, ncpu=4, nrhs=3)
run('p3d-10', M.poisson3d(10), 32, ncpu=2, nrhs=2)
run('rand', M.random_spd(500, 0.01, 1), 16, ncpu=3)
run('el3d-5', M.elasticity3d(5), 24, ncpu=2)
run('p2d-200', M.poisson2d(200), 256, ncpu=8)
run('p3d-30', M.poisson3d(30), 128, ncpu=8)
