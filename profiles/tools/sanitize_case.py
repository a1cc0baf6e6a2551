"""Small factor + solve runs for compute-sanitizer (racecheck / memcheck / synccheck):

  compute-sanitizer --tool racecheck python profiles/tools/sanitize_case.py p3d-12 el3d-6

Every solve path is exercised: the persistent pipelined kernels (nrhs 1 and 5), the level-set
launches (nrhs 9), forward-only + backward-only.  No torch import: host buffers through the C ABI.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spllt_b200 as sp          # noqa: E402
from spllt_b200 import matrices as M   # noqa: E402

CASES = {
    "p3d-12": (lambda: M.poisson3d(12), 48, 8),
    "el3d-6": (lambda: M.elasticity3d(6), 128, 4),
    "p2d-30": (lambda: M.poisson2d(30), 8, 4),
}


def run(name):
    mk, nb, ncpu = CASES[name]
    n, ptr, row, val = mk()
    s = sp.SpLLT(nb=nb, ncpu=ncpu)
    assert s.analyse(n, ptr, row) == 0
    for rep in range(2):       # second factorization = graph replay
        s.factor(val)
        s.wait()
    assert s.pivot_flag() == 0
    worst = 0.0
    for nrhs in (1, 5, 9):
        xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
        b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
        s.prepare_solve(nrhs)
        x = b.copy(order="F")
        assert s.solve(x, 0) == 0
        ok, err = sp.chkerr(n, ptr, row, val, x, b)
        assert ok == nrhs, err
        worst = max(worst, float(err.max()))
        x2 = b.copy(order="F")
        s.solve(x2, 1)
        s.solve(x2, 2)
        assert np.max(np.abs(x2 - x)) <= 1e-12 * np.abs(x).max()
    print("sanitize_case %s ok: n=%d nodes=%d worst bwd err %.2e" % (name, n, s.nnodes, worst), flush=True)


if __name__ == "__main__":
    for nme in (sys.argv[1:] or ["p3d-12", "el3d-6"]):
        run(nme)
