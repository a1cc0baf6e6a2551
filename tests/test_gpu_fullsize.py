"""Parity at the sizes BASELINE.json names (configs 2-5), through the C ABI against the CPU oracle.

Factor entries: every block column of the GPU factor against the oracle's (reference layout,
lower part), 1e-12 relative (north_star).  Solves: scaled backward error <= 1e-14
(src/utils_mod.F90:467) and the solution against the oracle's.  The oracle runs its OpenMP-task
driver on all host cores (64^3: < 1 s, 80^3: a few s, 100^3 / elasticity 60^3: tens of seconds).
"""
import os
import subprocess

import numpy as np
import pytest

import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle, chkerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACTOR_RTOL = 1e-12
BWD_TOL = 1e-14
CORES = os.cpu_count() or 1


def gpu_and_oracle(mat, nb, factor_oracle=True):
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=CORES)
    assert s.analyse(n, ptr, row) == 0
    assert s.factor(val) == 0
    s.wait()
    assert s.pivot_flag() == 0
    o = None
    if factor_oracle:
        sptr, sparent, rptr, rlist = s.symbolic()
        o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=CORES)
        o.factor(val, CORES)
    return s, o


def assert_factor_blockwise(s, o):
    """block column by block column (bounded host memory at 1e9 entries): lower part only -- the strict
    upper triangle of a diagonal tile is storage the reference never reads"""
    nb = s.options.nb
    nodes = s.nodes()
    sptr, sparent, rptr, rlist = s.symbolic()
    scale = 0.0
    worst_abs, worst_rel = 0.0, 0.0
    cols = []
    b = 0
    for k in range(s.nnodes):
        n = int(nodes[k, 1] - nodes[k, 0] + 1)
        m = int(rptr[k + 1] - rptr[k])
        for c0 in range(0, n, nb):
            b += 1
            cols.append((b, m - c0, min(nb, n - c0)))
    for b, h, w in cols:
        scale = max(scale, float(np.abs(o.lcol(b)).max()))
    for b, h, w in cols:
        g = s.lcol(b).reshape(h, w)
        r = o.lcol(b).reshape(h, w)
        if w > 1:
            iu = np.triu_indices(min(h, w), 1, w)
            g[iu] = 0.0
            r[iu] = 0.0
        d = np.abs(g - r)
        worst_abs = max(worst_abs, float(d.max()))
        big = np.abs(r) >= 1e-6 * scale
        if big.any():
            worst_rel = max(worst_rel, float(np.max(d[big] / np.abs(r[big]))))
    assert worst_abs <= FACTOR_RTOL * scale, (worst_abs, scale)
    assert worst_rel <= FACTOR_RTOL, worst_rel
    return worst_abs / scale, worst_rel


def rhs_for(mat, nrhs, seed=20261018):
    """odd columns: the reference's x = r convention (test/test_solve_phasis.F90:140-155); even: seeded random"""
    n, ptr, row, val = mat
    xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
    if nrhs > 1:
        xs[:, 1::2] = np.random.Generator(np.random.PCG64(seed)).standard_normal((n, xs[:, 1::2].shape[1]))
    return xs, np.asfortranarray(M.matvec(n, ptr, row, val, xs))


def check_solves(s, o, mat, nrhs_list, oracle_nrhs=()):
    n, ptr, row, val = mat
    for nrhs in nrhs_list:
        xs, b = rhs_for(mat, nrhs)
        s.prepare_solve(nrhs)
        x = b.copy(order="F")
        assert s.solve(x, 0) == 0
        ok, err = chkerr(n, ptr, row, val, x, b)
        assert ok == nrhs and err.max() <= BWD_TOL, (nrhs, err.max())
        assert np.max(np.abs(x - xs)) <= 1e-8 * np.abs(xs).max()
        if o is not None and nrhs in oracle_nrhs:
            o.prepare_solve(nrhs)
            xo = b.copy(order="F")
            o.solve(xo, 0)
            assert np.max(np.abs(x - xo)) <= 1e-10 * np.abs(xo).max()


def test_config2_p3d64_factor_entries_and_solve():
    """BASELINE configs[1]: 3D Poisson 64^3, nb = 512 -- the default persistent TMA tile path."""
    mat = M.poisson3d(64)
    s, o = gpu_and_oracle(mat, 512)
    assert_factor_blockwise(s, o)
    check_solves(s, o, mat, [1, 4], oracle_nrhs=(1, 4))


def test_config3_p3d100_factor_entries_and_solve():
    """BASELINE configs[2] (headline): 3D Poisson 100^3, nb = 768, factor entries vs the oracle."""
    mat = M.poisson3d(100)
    s, o = gpu_and_oracle(mat, 768)
    assert s.num_flops > 4e12
    assert_factor_blockwise(s, o)
    check_solves(s, o, mat, [1], oracle_nrhs=(1,))


@pytest.mark.parametrize("nb", [128, 512, 1024])
def test_config5_p3d80_solve_nrhs_nb_sweep(nb):
    """BASELINE configs[4]: 3D Poisson 80^3 solve phase, nrhs = 1 / 16 / 64, nb sweep 128-1024."""
    mat = M.poisson3d(80)
    s, o = gpu_and_oracle(mat, nb, factor_oracle=(nb == 512))
    if o is not None:
        assert_factor_blockwise(s, o)
    check_solves(s, o, mat, [1, 16, 64], oracle_nrhs=(16,))


def test_config4_elasticity60_backward_error_and_solution():
    """BASELINE configs[3]: 3D elasticity 27-point, 3 dof, 60^3 (n = 648 000, large fronts), nb = 768."""
    mat = M.elasticity3d(60)
    s, o = gpu_and_oracle(mat, 768)
    assert_factor_blockwise(s, o)
    check_solves(s, o, mat, [1, 16], oracle_nrhs=(1,))


def test_reanalyse_on_same_handles():
    """spllt_analyse may be called again on existing akeep / fkeep handles (the reference ABI allows
    it, interfaces/C/spllt_data_ciface.F90:151-163): every device buffer of the first matrix --
    solve work vectors, mailboxes, captured graphs -- must be dropped."""
    s = sp.SpLLT(nb=32, ncpu=2)
    for grid in (12, 17, 9):
        mat = M.poisson3d(grid)
        n, ptr, row, val = mat
        assert s.analyse(n, ptr, row) == 0
        s.factor(val)
        s.wait()
        assert s.pivot_flag() == 0
        check_solves(s, None, mat, [1, 3])


def test_factor_reports_not_pos_def_after_wait():
    """the first call on the handle after spllt_wait reports SPLLT_ERROR_NOT_POS_DEF (-20)"""
    n, ptr, row, good = M.poisson2d(12)
    val = good.copy()
    val[ptr[40] - 1] = -1.0
    s = sp.SpLLT(nb=16)
    s.analyse(n, ptr, row)
    s.factor(val)
    s.wait()
    s.prepare_solve_size(1)
    assert s.info.flag == -20
    assert s.pivot_flag() > 0       # 1-based pivot column (pivot order) of the failure
    # a following good factorization clears it
    s.factor(good)
    s.wait()
    s.prepare_solve_size(1)
    assert s.info.flag == 0


def test_c_example_compiles_and_runs(tmp_path):
    """examples/simple.c (the 3x3 system of the reference's example/C/simple.c:25-52) against
    include/spllt_iface.h and libspllt_b200.so only."""
    exe = str(tmp_path / "simple")
    libdir = os.path.join(ROOT, "spllt_b200")
    cmd = ["gcc", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "simple.c"), "-L" + libdir,
           "-lspllt_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "expected 1.5 2 1.5" in r.stdout and "ok for   1/  1" in r.stdout
