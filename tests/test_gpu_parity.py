"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes mirror), against
the CPU oracle on identical inputs -- factor entries, solutions, the reference's acceptance
gate -- plus the committed golden fixtures and size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star): factor entries within 1e-12 relative; scaled backward
error <= 1e-14.  "Relative" for factor entries is taken entry-wise for every entry that is not
negligible (|L_ij| >= 1e-6 max|L|), and as |diff| <= 1e-12 max|L| for the rest (entries that
are the result of cancellation have no meaningful entry-wise relative error).
The strict upper triangle of a diagonal tile is storage the reference never reads (dpotrf 'U'
on the transposed view, src/spllt_kernels_mod.F90:1179) and is excluded.
"""
import glob
import os

import numpy as np
import pytest

import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle, chkerr
from tests.cases import SMALL, MEDIUM, ids

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FACTOR_RTOL = 1e-12
BWD_TOL = 1e-14


def lower_mask(s):
    """True for the entries of the reference layout that belong to L (drops the strict upper
    triangle of diagonal tiles)."""
    out = []
    nb = s.options.nb if s.options.nb >= 1 else 256
    nodes = s.nodes()
    sptr, sparent, rptr, rlist = s.symbolic()
    for k in range(s.nnodes):
        n = int(nodes[k, 1] - nodes[k, 0] + 1)
        m = int(rptr[k + 1] - rptr[k])
        for c0 in range(0, n, nb):
            w, h = min(nb, n - c0), m - c0
            out.append(np.tril(np.ones((h, w), bool)).ravel())
    return np.concatenate(out) if out else np.zeros(0, bool)


def assert_factor_close(got, ref, mask):
    got, ref = got[mask], ref[mask]
    scale = np.abs(ref).max()
    d = np.abs(got - ref)
    assert d.max() <= FACTOR_RTOL * scale
    big = np.abs(ref) >= 1e-6 * scale
    assert np.max(d[big] / np.abs(ref[big])) <= FACTOR_RTOL


def both(case, nthreads=1):
    name, mk, nb, ncpu, prune = case
    n, ptr, row, val = mk()
    s = sp.SpLLT(nb=nb, ncpu=ncpu, prune_tree=prune)
    assert s.analyse(n, ptr, row) == 0
    sptr, sparent, rptr, rlist = s.symbolic()
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu, prune=prune)
    o.factor(val, nthreads)
    assert s.factor(val) == 0
    s.wait()
    assert s.pivot_flag() == 0
    return s, o, (n, ptr, row, val)


def rhs_for(mat, nrhs, seed=None):
    n, ptr, row, val = mat
    if seed is None:   # reference convention, test/test_solve_phasis.F90:140-155
        xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
    else:
        xs = np.asfortranarray(np.random.default_rng(seed).standard_normal((n, nrhs)))
    return xs, np.asfortranarray(M.matvec(n, ptr, row, val, xs))


@pytest.mark.parametrize("case", SMALL + MEDIUM, ids=ids(SMALL + MEDIUM))
def test_factor_entries_vs_oracle(case):
    s, o, mat = both(case)
    assert_factor_close(s.factor_entries(), o.factor_entries(), lower_mask(s))


@pytest.mark.parametrize("case", SMALL + MEDIUM, ids=ids(SMALL + MEDIUM))
@pytest.mark.parametrize("nrhs", [1, 3, 16])
def test_solve_vs_oracle_and_gate(case, nrhs):
    s, o, mat = both(case)
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, nrhs)
    s.prepare_solve(nrhs)
    x = b.copy(order="F")
    assert s.solve(x, 0) == 0
    ok, err = chkerr(n, ptr, row, val, x, b)          # the oracle's check_backward_error
    assert ok == nrhs and err.max() <= BWD_TOL
    o.prepare_solve(nrhs)
    xo = b.copy(order="F")
    o.solve(xo, 0)
    assert np.max(np.abs(x - xo)) <= 1e-10 * np.abs(xo).max()


@pytest.mark.parametrize("case", [SMALL[4], SMALL[8], SMALL[10], MEDIUM[1]],
                         ids=ids([SMALL[4], SMALL[8], SMALL[10], MEDIUM[1]]))
@pytest.mark.parametrize("nrhs", [1, 5])
def test_job1_then_job2_equals_job0(case, nrhs):
    """test/test_solve_phasis.F90:253-262: forward (job 1) then backward (job 2)."""
    s, o, mat = both(case)
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, nrhs, seed=20261018)
    s.prepare_solve(nrhs)
    x0 = b.copy(order="F")
    s.solve(x0, 0)
    x1 = b.copy(order="F")
    assert s.solve(x1, 1) == 0
    assert np.array_equal(x1, b)          # job 1 leaves x untouched; the result lives in y
    # forward result vs the oracle's y (pivot order; reference L1 row-block layout)
    o.prepare_solve(nrhs)
    xo = b.copy(order="F")
    o.solve(xo, 1)
    yo = o.y(nrhs)
    assert np.max(np.abs(s.y[:n * nrhs] - yo)) <= 1e-10 * np.abs(yo).max()
    assert s.solve(x1, 2) == 0
    assert np.max(np.abs(x1 - x0)) <= 1e-13 * np.abs(x0).max()
    ok, err = chkerr(n, ptr, row, val, x1, b)
    assert ok == nrhs


def test_invalid_job_is_rejected():
    s, o, mat = both(SMALL[3])
    n = mat[0]
    x = np.ones(n)
    assert s.solve(x, 6) == -10 and np.all(x == 1.0)      # example/C/simple.c:69 passes 6
    assert s.solve(x, -1) == -10


def test_known_answer_simple_c():
    s, o, mat = both(SMALL[0])
    x = np.ones(3)
    s.prepare_solve(1)
    s.solve(x, 0)
    assert np.allclose(x, [1.5, 2.0, 1.5], rtol=0, atol=2e-15)


def test_async_worker_form_and_refactor():
    """spllt_solve_worker + spllt_wait; a second spllt_factor with new values reuses the analysis."""
    s, o, mat = both(SMALL[9])
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, 2)
    x = b.copy(order="F")
    s.solve_worker(x, 0)
    s.wait()
    assert chkerr(n, ptr, row, val, x, b)[0] == 2
    val2 = val * 3.0
    s.factor(val2)
    s.wait()
    x = b.copy(order="F")
    s.solve(x, 0)
    assert chkerr(n, ptr, row, val2, x, b)[0] == 2


def test_not_positive_definite_is_reported():
    n, ptr, row, val = M.poisson2d(8)
    val = val.copy()
    val[ptr[10] - 1] = -4.0          # a negative diagonal entry
    s = sp.SpLLT(nb=16)
    s.analyse(n, ptr, row)
    s.factor(val)
    s.wait()
    assert s.pivot_flag() > 0
    x = np.ones(n)
    assert s.solve(x, 0) == -20


def test_spllt_all_chain():
    import ctypes as C
    n, ptr, row, val = M.poisson2d(15)
    xs, b = rhs_for((n, ptr, row, val), 2)
    x = b.copy(order="F")
    L = sp.lib()
    ak, fk = C.c_void_p(None), C.c_void_p(None)
    opt, info = sp.Options(nb=8), sp.Inform()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.spllt_all(C.byref(ak), C.byref(fk), C.byref(opt), n, val.size, 2, 8, ptr.ctypes.data_as(ip),
                row.ctypes.data_as(ip), val.ctypes.data_as(dp), x.ctypes.data_as(dp), b.ctypes.data_as(dp),
                C.byref(info))
    assert info.flag == 0 and info.num_nodes > 0
    assert chkerr(n, ptr, row, val, x, b)[0] == 2
    st = C.c_int(0)
    L.spllt_deallocate_akeep(C.byref(ak), C.byref(st))
    L.spllt_deallocate_fkeep(C.byref(fk), C.byref(st))
    assert ak.value is None and fk.value is None


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_golden_fixtures(path):
    g = np.load(path)
    n, ptr, row, val = int(g["n"]), g["ptr"], g["row"], g["val"]
    s = sp.SpLLT(nb=int(g["nb"]), ncpu=int(g["ncpu"]))
    s.analyse(n, ptr, row)
    assert np.array_equal(g["order"], s.order[:n]) and np.array_equal(g["blocks"], s.blocks())
    s.factor(val)
    s.wait()
    assert_factor_close(s.factor_entries(), g["factor"], lower_mask(s))
    x = np.asfortranarray(g["rhs"].copy())
    s.prepare_solve(x.shape[1])
    s.solve(x, 0)
    assert np.max(np.abs(x - g["x"])) <= 1e-12 * np.abs(g["x"]).max()


@pytest.mark.parametrize("nb", [4, 5, 16, 64, 100, 256])
def test_nb_sweep(nb):
    """scripts/stress_test.sh sweeps nb; odd nb exercises the unaligned (8-byte cp.async) tile loads."""
    case = ("p3d-9-nb%d" % nb, lambda: M.poisson3d(9), nb, 2, 1)
    s, o, mat = both(case)
    assert_factor_close(s.factor_entries(), o.factor_entries(), lower_mask(s))
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, 2)
    x = b.copy(order="F")
    s.prepare_solve(2)
    s.solve(x, 0)
    assert chkerr(n, ptr, row, val, x, b)[0] == 2


@pytest.mark.parametrize("case", [SMALL[9], SMALL[11], SMALL[13], MEDIUM[1], MEDIUM[3]],
                         ids=ids([SMALL[9], SMALL[11], SMALL[13], MEDIUM[1], MEDIUM[3]]))
@pytest.mark.parametrize("tma", [0, 1])
def test_large_tile_kernels_forced(case, tma, monkeypatch):
    """Forces the 128 x 128 tile kernels (persistent TMA / cp.async variants) onto small problems so
    ragged tiles, K tails and the scatter epilogue are all checked against the oracle."""
    monkeypatch.setenv("SPLLT_B200_TILE_WAVE", "1")
    monkeypatch.setenv("SPLLT_B200_TILE_L_MIN", "32")
    if not tma:
        monkeypatch.setenv("SPLLT_B200_NO_TMA", "1")
    s, o, mat = both(case)
    assert_factor_close(s.factor_entries(), o.factor_entries(), lower_mask(s))


def test_full_size_properties_p3d64():
    """BASELINE config 2 (3D Poisson 64^3, nb = 512): too large for the oracle to finish in
    seconds, so parity is checked through size-independent properties: the acceptance gate,
    linearity of the solve, and A x = b residuals for a seeded random right-hand side."""
    mat = M.poisson3d(64)
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=512)
    s.analyse(n, ptr, row)
    assert s.num_flops > 1e11
    s.factor(val)
    s.wait()
    assert s.pivot_flag() == 0
    xs, b = rhs_for(mat, 4, seed=20261018)
    s.prepare_solve(4)
    x = b.copy(order="F")
    s.solve(x, 0)
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == 4 and err.max() <= BWD_TOL
    assert np.max(np.abs(x - xs)) <= 1e-9 * np.abs(xs).max()
    # linearity: solve(b0 + 2 b1) == x0 + 2 x1
    s.prepare_solve(1)
    c = np.asfortranarray(b[:, 0] + 2.0 * b[:, 1])
    s.solve(c, 0)
    assert np.max(np.abs(c - (x[:, 0] + 2.0 * x[:, 1]))) <= 1e-10 * np.abs(c).max()


def test_cli_driver(capsys):
    """drivers/spllt_test.F90 role: flags, timing lines and the backward-error report."""
    from spllt_b200 import driver
    rc = driver.main(["--mat", "poisson3d:12", "--nb", "32", "--nrhs", "3", "--ncpu", "2"])
    out = capsys.readouterr().out
    assert rc == 0 and "Factor took" in out and "ok for 3/3" in out


@pytest.mark.parametrize("nrhs", [2, 7, 8, 9, 33, 64])
def test_nrhs_sweep(nrhs):
    """scripts/stress_test.sh sweeps nrhs in {1..10, 16, 32, 64, 128}."""
    s, o, mat = both(MEDIUM[1])
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, nrhs, seed=7)
    s.prepare_solve(nrhs)
    x = b.copy(order="F")
    assert s.solve(x, 0) == 0
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == nrhs and err.max() <= BWD_TOL


def test_elasticity_config_small():
    """BASELINE config 4 (27-point, 3 dof) at a size the oracle handles: large fronts relative to n."""
    case = ("el3d-14-nb256", lambda: M.elasticity3d(14), 256, 4, 1)
    s, o, mat = both(case)
    assert_factor_close(s.factor_entries(), o.factor_entries(), lower_mask(s))
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, 2)
    x = b.copy(order="F")
    s.prepare_solve(2)
    s.solve(x, 0)
    assert chkerr(n, ptr, row, val, x, b)[0] == 2


@pytest.mark.parametrize("env", [
    {"SPLLT_B200_GRAPH": "0"},
    {"SPLLT_B200_NO_OVERLAP": "1"},
    {"SPLLT_B200_TILE_N": "128", "SPLLT_B200_TILE_WAVE": "1", "SPLLT_B200_TILE_L_MIN": "32"},
    {"SPLLT_B200_DEFER": "1", "SPLLT_B200_TILE_WAVE": "1", "SPLLT_B200_TILE_L_MIN": "32"},
    {"SPLLT_B200_EXCL": "1", "SPLLT_B200_EXCL_MIN": "1", "SPLLT_B200_TILE_WAVE": "1", "SPLLT_B200_TILE_L_MIN": "32"},
], ids=["nograph", "nooverlap", "tile128", "deferred-bg-stream", "exclusive-no-atomics"])
@pytest.mark.parametrize("case", [SMALL[11], MEDIUM[1]], ids=ids([SMALL[11], MEDIUM[1]]))
def test_schedule_variants(case, env, monkeypatch):
    """Every opt-in / fallback execution mode of the factorization schedule produces the same factor."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    s, o, mat = both(case)
    assert_factor_close(s.factor_entries(), o.factor_entries(), lower_mask(s))
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, 2)
    x = b.copy(order="F")
    s.prepare_solve(2)
    s.solve(x, 0)
    assert chkerr(n, ptr, row, val, x, b)[0] == 2


@pytest.mark.parametrize("env", [
    {},
    {"SPLLT_B200_PIPE_MAX_NRHS": "8"},
    {"SPLLT_B200_SOLVE_LEVELSET": "1"},
    {"SPLLT_B200_SOLVE_CUT": "2", "SPLLT_B200_PIPE_MAX_NRHS": "8"},
    {"SPLLT_B200_SOLVE_CUT": "5", "SPLLT_B200_GRAPH": "0", "SPLLT_B200_PIPE_MAX_NRHS": "8"},
    {"SPLLT_B200_PIPE_MAX_NRHS": "8", "SPLLT_B200_PIPE_MODE": "96"},
    {"SPLLT_B200_PIPE_MAX_NRHS": "8", "SPLLT_B200_PIPE_ORDER_EST": "0"},
    {"SPLLT_B200_PIPE_MAX_NRHS": "8", "SPLLT_B200_PIPE_ORDER_EST": "0", "SPLLT_B200_PIPE_BWD_EARLY": "0"},
], ids=["auto", "pipelined", "levelset", "cut2", "cut5-nograph", "pipelined-flags-only",
        "pipelined-backward-readiness-order", "pipelined-depth-node-order"])
@pytest.mark.parametrize("case", [SMALL[7], SMALL[11], MEDIUM[1], MEDIUM[2], MEDIUM[3]],
                         ids=ids([SMALL[7], SMALL[11], MEDIUM[1], MEDIUM[2], MEDIUM[3]]))
@pytest.mark.parametrize("nrhs", [1, 6])
def test_solve_variants(case, env, nrhs, monkeypatch):
    """The path chosen from the matrix structure, the persistent pipelined solve forced, the
    level-set launches, hybrids of the two and the pipelined solve without the mailbox all give
    the oracle's solution; forward-only + backward-only equals the full solve."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    s, o, mat = both(case)
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, nrhs, seed=3)
    s.prepare_solve(nrhs)
    o.prepare_solve(nrhs)
    xo = b.copy(order="F")
    o.solve(xo, 0)
    for rep in range(3):          # replays of the captured graph must reset flags and counters
        x = b.copy(order="F")
        assert s.solve(x, 0) == 0
        ok, err = chkerr(n, ptr, row, val, x, b)
        assert ok == nrhs and err.max() <= BWD_TOL
        assert np.max(np.abs(x - xo)) <= 1e-10 * np.abs(xo).max()
    x2 = b.copy(order="F")
    assert s.solve(x2, 1) == 0
    assert s.solve(x2, 2) == 0
    assert np.max(np.abs(x2 - xo)) <= 1e-10 * np.abs(xo).max()


@pytest.mark.parametrize("case", [SMALL[11], MEDIUM[1], MEDIUM[3]], ids=ids([SMALL[11], MEDIUM[1], MEDIUM[3]]))
@pytest.mark.parametrize("nrhs", [9, 19])
def test_pipelined_solve_many_rhs(case, nrhs, monkeypatch):
    """By default more than 8 right-hand sides go to the level-set kernels; forced through the
    persistent kernels (several passes of 4 right-hand sides, the last one partial) the result
    is the same."""
    monkeypatch.setenv("SPLLT_B200_PIPE_MAX_NRHS", "64")
    s, o, mat = both(case)
    n, ptr, row, val = mat
    xs, b = rhs_for(mat, nrhs, seed=5)
    s.prepare_solve(nrhs)
    o.prepare_solve(nrhs)
    xo = b.copy(order="F")
    o.solve(xo, 0)
    x = b.copy(order="F")
    assert s.solve(x, 0) == 0
    ok, err = chkerr(n, ptr, row, val, x, b)
    assert ok == nrhs and err.max() <= BWD_TOL
    assert np.max(np.abs(x - xo)) <= 1e-10 * np.abs(xo).max()


def test_unsorted_input_columns():
    """Entries of a column may come in any order; the A -> L map follows the input order."""
    n, ptr, row, val = M.poisson3d(8)
    rng = np.random.default_rng(1)
    row2, val2 = row.copy(), val * (1.0 + 0.01 * rng.standard_normal(val.size))
    diag = row == np.repeat(np.arange(1, n + 1), np.diff(ptr))
    val2[diag] = np.abs(val2[diag]) + 1.0
    val3 = val2.copy()
    for j in range(n):
        a, b = ptr[j] - 1, ptr[j + 1] - 1
        p = rng.permutation(b - a)
        row2[a:b], val3[a:b] = row[a:b][p], val2[a:b][p]
    xs = np.ones(n)
    bvec = M.matvec(n, ptr, row, val2, xs)
    for r_, v_ in ((row, val2), (row2, val3)):
        s = sp.SpLLT(nb=16)
        s.analyse(n, ptr, r_)
        s.factor(v_)
        s.wait()
        x = bvec.copy()
        s.prepare_solve(1)
        s.solve(x, 0)
        assert chkerr(n, ptr, row, val2, x, bvec)[0] == 1
