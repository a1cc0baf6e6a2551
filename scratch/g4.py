import sys, numpy as np
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
N = int(sys.argv[1]); nb = int(sys.argv[2])
mat = M.poisson3d(N); n, ptr, row, val = mat
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
for rep in range(3):
    s.factor(val); s.wait()
    rng = np.random.default_rng(20261018)
    xs = np.asfortranarray(rng.standard_normal((n, 4)))
    b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
    s.prepare_solve(4); x = b.copy(order='F'); s.solve(x, 0)
    ok, err = sp.chkerr(n, ptr, row, val, x, b)
    print('rep', rep, 'pivot', s.pivot_flag(), 'ok', ok, 'err', err, 'fwd', np.abs(x-xs).max())
