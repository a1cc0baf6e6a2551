"""Shared test-case table: (name, matrix factory, nb, ncpu, prune)."""
import numpy as np

from spllt_b200 import matrices as M

SMALL = [
    ("tri3", lambda: M.tridiag3(), 4, 1, 1),
    ("n1", lambda: (1, np.array([1, 2], np.int32), np.array([1], np.int32), np.array([4.0])), 16, 1, 1),
    ("diag7", lambda: (7, np.arange(1, 9, dtype=np.int32), np.arange(1, 8, dtype=np.int32), np.arange(1.0, 8.0)),
     4, 2, 1),
    ("p2d-20-nb16", lambda: M.poisson2d(20), 16, 1, 1),
    ("p2d-30-nb8", lambda: M.poisson2d(30), 8, 4, 1),
    ("p2d-17x9-nb5", lambda: M.poisson2d(17, 9), 5, 3, 1),
    ("p2d-31-nb7-noprune", lambda: M.poisson2d(31), 7, 2, 0),
    ("p3d-10-nb32", lambda: M.poisson3d(10), 32, 2, 1),
    ("p3d-12-nb48", lambda: M.poisson3d(12), 48, 8, 1),
    ("p3d-9x7x5-nb33", lambda: M.poisson3d(9, 7, 5), 33, 1, 1),
    ("rand500", lambda: M.random_spd(500, 0.01, 1), 16, 3, 1),
    ("rand300-dense", lambda: M.random_spd(300, 0.3, 2), 64, 2, 1),
    ("el3d-5-nb24", lambda: M.elasticity3d(5), 24, 2, 1),
    ("el3d-6-nb128", lambda: M.elasticity3d(6), 128, 4, 1),
]

MEDIUM = [
    ("p2d-200-nb256", lambda: M.poisson2d(200), 256, 8, 1),   # BASELINE config 1
    ("p3d-30-nb128", lambda: M.poisson3d(30), 128, 8, 1),
    ("p3d-40-nb256", lambda: M.poisson3d(40), 256, 1, 1),
    ("el3d-12-nb192", lambda: M.elasticity3d(12), 192, 8, 1),
]


def ids(cases):
    return [c[0] for c in cases]
