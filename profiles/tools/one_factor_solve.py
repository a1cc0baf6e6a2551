"""Minimal program for ncu captures: analyse, N factorizations, M solves of one workload through the
C ABI (host buffers, no torch).  usage: one_factor_solve.py <workload> [nfactor] [nsolve] [nrhs]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spllt_b200 as sp   # noqa: E402
import bench              # noqa: E402

wl = sys.argv[1]
nfac = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nsol = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nrhs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
(n, ptr, row, val), nb, desc = bench.make_matrix(wl)
s = sp.SpLLT(nb=nb)
s.analyse(n, ptr, row)
for _ in range(nfac):
    s.factor(val)
    s.wait()
assert s.pivot_flag() == 0
M = bench.matrices_module()
xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
s.prepare_solve(nrhs)
err = np.zeros(1)
for _ in range(nsol):
    x = b.copy(order="F")
    s.solve(x, 0)
if nsol:
    ok, err = sp.chkerr(n, ptr, row, val, x, b)
print(desc, "factor launches", s.L.spllt_b200_factor_launches(s.fkeep), "bwd err %.2e" % err.max())
