"""Synthetic SPD test matrices of BASELINE.json's configs.

All generators return (n, ptr, row, val): lower triangle, CSC, 1-based int32 indices -- the
convention of the reference's C interface (example/C/simple.c:38-39).  No randomness.
Names follow the reference's benchmark scripts (aux/run_tests_poisson3d.sh:9-10).
"""
import numpy as np


def _lower_csc(n, rows, cols, vals):
    """COO (0-based, lower triangle, no duplicates) -> 1-based lower CSC sorted by (col, row)."""
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, cols + 1, 1)
    ptr = np.cumsum(ptr) + 1
    return n, ptr.astype(np.int32), (rows + 1).astype(np.int32), vals.astype(np.float64)


def _stencil(dims, offsets, diag, off):
    """Scalar stencil on a regular grid, lexicographic numbering (x fastest)."""
    dims = tuple(int(d) for d in dims)
    n = int(np.prod(dims))
    grids = np.meshgrid(*[np.arange(d) for d in dims[::-1]], indexing="ij")  # slowest first
    coords = [g.ravel() for g in grids][::-1]  # coords[0] = fastest axis
    strides = np.cumprod((1,) + dims[:-1])
    idx = sum(c * s for c, s in zip(coords, strides))
    rows = [idx]
    cols = [idx]
    vals = [np.full(n, float(diag))]
    for o in offsets:  # only offsets that lead to a larger index (lower triangle, col < row)
        ok = np.ones(n, dtype=bool)
        for c, d, oo in zip(coords, dims, o):
            ok &= (c + oo >= 0) & (c + oo < d)
        tgt = idx + sum(oo * s for oo, s in zip(o, strides))
        rows.append(tgt[ok])
        cols.append(idx[ok])
        vals.append(np.full(int(ok.sum()), float(off)))
    return _lower_csc(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def poisson2d(nx, ny=None):
    """5-point Laplacian, diag 4, off-diagonal -1 (BASELINE config 1: 200 x 200)."""
    ny = nx if ny is None else ny
    return _stencil((nx, ny), [(1, 0), (0, 1)], 4.0, -1.0)


def poisson3d(nx, ny=None, nz=None):
    """7-point Laplacian, diag 6, off-diagonal -1 (BASELINE configs 2, 3, 5)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return _stencil((nx, ny, nz), [(1, 0, 0), (0, 1, 0), (0, 0, 1)], 6.0, -1.0)


def elasticity3d(nx, ny=None, nz=None, dof=3):
    """27-point stencil x `dof` unknowns per grid node (BASELINE config 4: 60^3 x 3).

    Every pair of stencil neighbours is coupled by the full dof x dof block
    B = -(1/26) (I + 0.1 * ones); the diagonal block is D = d I + 0.05 (ones - I) with d = 2.0,
    which makes every row strictly diagonally dominant (26 * (1.3/26) + 0.1 = 1.4 < 2).
    """
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    dims = (nx, ny, nz)
    offs = [(a, b, c) for c in (-1, 0, 1) for b in (-1, 0, 1) for a in (-1, 0, 1)
            if (c, b, a) > (0, 0, 0)]  # the 13 neighbours with a larger lexicographic index
    nn = nx * ny * nz
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    coords = (xx.ravel(), yy.ravel(), zz.ravel())
    strides = (1, nx, nx * ny)
    idx = sum(c * s for c, s in zip(coords, strides))
    rows, cols, vals = [], [], []
    # diagonal blocks (lower triangle of D)
    for a in range(dof):
        for b in range(a + 1):
            rows.append(idx * dof + a)
            cols.append(idx * dof + b)
            vals.append(np.full(nn, 2.0 if a == b else 0.05))
    blk = -(1.0 / 26.0) * (np.eye(dof) + 0.1 * np.ones((dof, dof)))
    for o in offs:
        ok = np.ones(nn, dtype=bool)
        for c, d, oo in zip(coords, dims, o):
            ok &= (c + oo >= 0) & (c + oo < d)
        src = idx[ok]
        tgt = src + sum(oo * s for oo, s in zip(o, strides))
        for a in range(dof):
            for b in range(dof):
                rows.append(tgt * dof + a)
                cols.append(src * dof + b)
                vals.append(np.full(src.size, blk[a, b]))
    return _lower_csc(nn * dof, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def tridiag3():
    """The 3 x 3 known-answer system of example/C/simple.c:25-52: x = [1.5, 2, 1.5] for b = 1."""
    ptr = np.array([1, 3, 5, 6], dtype=np.int32)
    row = np.array([1, 2, 2, 3, 3], dtype=np.int32)
    val = np.array([2.0, -1.0, 2.0, -1.0, 2.0])
    return 3, ptr, row, val


def random_spd(n, density, seed):
    """Random sparse SPD matrix (strictly diagonally dominant), seeded."""
    rng = np.random.default_rng(seed)
    nnz = max(1, int(density * n * n / 2))
    r = rng.integers(0, n, nnz)
    c = rng.integers(0, n, nnz)
    keep = r != c
    r, c = np.maximum(r[keep], c[keep]), np.minimum(r[keep], c[keep])
    key = np.unique(c.astype(np.int64) * n + r)
    c, r = (key // n).astype(np.int64), (key % n).astype(np.int64)
    v = rng.uniform(-1.0, 1.0, r.size)
    rowsum = np.zeros(n)
    np.add.at(rowsum, r, np.abs(v))
    np.add.at(rowsum, c, np.abs(v))
    d = rowsum + 1.0 + rng.uniform(0, 1, n)
    ar = np.arange(n)
    return _lower_csc(n, np.concatenate([ar, r]), np.concatenate([ar, c]), np.concatenate([d, v]))


def matvec(n, ptr, row, val, x):
    """y = A x for the symmetric matrix given by its lower triangle; x is n or n x nrhs (column-major)."""
    x2 = x.reshape(n, -1, order="F") if x.ndim == 1 else x
    cols = np.repeat(np.arange(n), np.diff(ptr))
    rows = row.astype(np.int64) - 1
    y = np.zeros_like(x2, dtype=np.float64)
    np.add.at(y, rows, val[:, None] * x2[cols])
    offd = rows != cols
    np.add.at(y, cols[offd], val[offd, None] * x2[rows[offd]])
    return y.reshape(x.shape, order="F") if x.ndim == 1 else y


def to_dense(n, ptr, row, val):
    a = np.zeros((n, n))
    cols = np.repeat(np.arange(n), np.diff(ptr))
    rows = row.astype(np.int64) - 1
    a[rows, cols] = val
    a[cols, rows] = val
    return a
