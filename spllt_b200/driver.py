"""Command-line driver: the role of the reference's drivers/spllt_test.F90 (same flags as
spllt_parse_args, src/spllt_mod.F90:328-421: --nb --ncpu --nemin --nrhs --mat / --mm), on top of
the C ABI.  Prints the reference's timing lines ("Factor took", "bwd error scaled").

  python -m spllt_b200.driver --mat poisson3d:64 --nb 512 --nrhs 4
  python -m spllt_b200.driver --mm matrix.mtx --nb 256
"""
import argparse
import time

import numpy as np

from . import matrices as M
from .api import SpLLT, chkerr


def read_matrix_market(path):
    """Symmetric Matrix Market file -> lower CSC, 1-based (the COO -> CSC step of
    src/spllt_mod.F90:426-620).  Duplicate entries are summed."""
    import scipy.io
    import scipy.sparse as sps
    a = scipy.io.mmread(path)
    a = sps.tril(sps.csc_matrix(a), format="csc")
    a.sum_duplicates()
    a.sort_indices()
    n = a.shape[0]
    return n, (a.indptr + 1).astype(np.int32), (a.indices + 1).astype(np.int32), a.data.astype(np.float64)


def make(spec):
    name, _, arg = spec.partition(":")
    dims = [int(x) for x in arg.split("x")] if arg else [32]
    return getattr(M, name)(*dims)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--ncpu", type=int, default=1)
    ap.add_argument("--nemin", type=int, default=32)
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--no-prune-tree", action="store_true")
    ap.add_argument("--mat", default=None, help="generator[:dims], e.g. poisson2d:200, poisson3d:64, elasticity3d:20")
    ap.add_argument("--mm", default=None, help="Matrix Market file (symmetric positive definite)")
    ap.add_argument("--repeat", type=int, default=1)
    args = ap.parse_args(argv)
    n, ptr, row, val = read_matrix_market(args.mm) if args.mm else make(args.mat or "poisson3d:32")
    print(" [>] n = %d, nnz(lower) = %d" % (n, val.size))
    s = SpLLT(nb=args.nb, ncpu=args.ncpu, nemin=args.nemin, prune_tree=0 if args.no_prune_tree else 1)
    t = time.perf_counter()
    s.analyse(n, ptr, row)
    print(" [>] [analysis] took %.3f s; nodes %d, nnz(L) %d, flops %.3e" %
          (time.perf_counter() - t, s.nnodes, s.num_factor, s.num_flops))
    for _ in range(args.repeat):
        t = time.perf_counter()
        s.factor(val)
        s.wait()
        tf = time.perf_counter() - t
    print(" Factor took %.4f s  (%.1f GFLOP/s incl. H2D of val)" % (tf, s.num_flops / tf / 1e9))
    if s.pivot_flag():
        print(" Error: matrix is not positive definite (pivot %d)" % s.pivot_flag())
        return 1
    xs = np.asfortranarray(np.tile(np.arange(1.0, args.nrhs + 1), (n, 1)))    # test/test_solve_phasis.F90:140-155
    rhs = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
    x = rhs.copy(order="F")
    s.prepare_solve(args.nrhs)
    t = time.perf_counter()
    s.solve(x, 0)
    print(" Solve took %.4f s" % (time.perf_counter() - t))
    ok, err = chkerr(n, ptr, row, val, x, rhs)
    print(" bwd error scaled = %.3e ; Backward error... ok for %d/%d" % (err.max(), ok, args.nrhs))
    print(" fwd error || ||_inf = %.3e" % np.abs(x - xs).max())
    return 0 if ok == args.nrhs else 2


if __name__ == "__main__":
    raise SystemExit(main())
