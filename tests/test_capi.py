"""The C-ABI shared library loads and exports every symbol include/*.h declares; host-only entry
points behave like the reference's (no compute calls here: there is no GPU on the CPU box)."""
import ctypes as C
import os
import re

import numpy as np

import spllt_b200 as sp
from spllt_b200 import matrices as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("spllt_iface.h", "spllt_b200.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(spllt_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = sp.load_library()
    names = declared_symbols()
    assert len(names) >= 40
    for nme in names:
        assert hasattr(L, nme), nme
    # and the ctypes mirror binds all of them
    assert set(names) <= set(L._signatures), set(names) - set(L._signatures)


def test_struct_layouts_match_reference_header():
    assert C.sizeof(sp.Options) == 14 * 4      # include/spllt_iface.h:14-31
    assert C.sizeof(sp.Inform) == 6 * 4        # :49-57
    assert [f[0] for f in sp.Options._fields_][:4] == ["print_level", "nrhs", "ncpu", "nb"]


def test_analyse_is_host_only_and_fills_inform():
    n, ptr, row, val = M.poisson2d(10)
    s = sp.SpLLT(nb=8)
    assert s.analyse(n, ptr, row) == 0
    assert s.info.num_nodes == s.nnodes > 0
    assert s.info.num_factor == s.num_factor and s.info.num_flops == s.num_flops
    assert sorted(s.order[:n]) == list(range(1, n + 1))
    ws = s.prepare_solve_size(3)
    sptr, sparent, rptr, rlist = s.symbolic()
    assert ws == 3 * int(np.sum(np.diff(rptr) - np.diff(sptr)))
    size = C.c_long(0)
    s.L.spllt_solve_workspace_size(s.fkeep, 2, 3, C.byref(size))       # src/spllt_data_mod.F90:655
    assert size.value == n * 3 + (s.L.spllt_b200_maxmn(s.akeep) + n) * 3 * 2
    s.free()
    assert s.akeep.value is None and s.fkeep.value is None


def test_invalid_job_needs_no_gpu():
    n, ptr, row, val = M.poisson2d(6)
    s = sp.SpLLT(nb=8)
    s.analyse(n, ptr, row)
    x = np.ones(n)
    assert s.solve(x, 7) == -10        # SPLLT_WARNING_PARAM_VALUE, src/spllt_solve_mod.F90:216-220
    assert np.all(x == 1.0)


def test_task_manager_handles():
    L = sp.lib()
    tm, st = C.c_void_p(None), C.c_int(5)
    L.spllt_task_manager_init(C.byref(tm))
    assert tm.value
    L.spllt_task_manager_deallocate(C.byref(tm), C.byref(st))
    assert tm.value is None and st.value == 0


def test_user_and_natural_ordering():
    n, ptr, row, val = M.poisson2d(7)
    s = sp.SpLLT(nb=8)
    s.analyse(n, ptr, row, ordering=sp.ORDER_NATURAL)
    assert np.array_equal(np.sort(s.order[:n]), np.arange(1, n + 1))
    perm = np.random.default_rng(3).permutation(n).astype(np.int32) + 1
    s2 = sp.SpLLT(nb=8)
    s2.analyse(n, ptr, row, ordering=sp.ORDER_USER, order=perm)
    assert sorted(s2.order[:n]) == list(range(1, n + 1))


def test_matrix_market_reader(tmp_path):
    """COO -> lower CSC conversion of the driver (role of src/spllt_mod.F90:426-620)."""
    import scipy.io
    import scipy.sparse as sps
    from spllt_b200.driver import read_matrix_market
    n, ptr, row, val = M.poisson2d(5)
    a = M.to_dense(n, ptr, row, val)
    p = tmp_path / "a.mtx"
    scipy.io.mmwrite(str(p), sps.coo_matrix(a), symmetry="symmetric")
    n2, ptr2, row2, val2 = read_matrix_market(str(p))
    assert n2 == n and np.array_equal(ptr2, ptr) and np.array_equal(row2, row) and np.allclose(val2, val)


def test_empty_matrix_and_unsorted_rows():
    s = sp.SpLLT(nb=8)
    assert s.analyse(0, np.array([1], np.int32), np.array([], np.int32)) == 0
    assert s.nnodes == 0 and s.num_flops == 0 and s.prepare_solve_size(2) == 0
    # row indices need not be sorted inside a column (the reference goes through spllt_make_map)
    n, ptr, row, val = M.poisson2d(6)
    rng = np.random.default_rng(0)
    row2 = row.copy()
    for j in range(n):
        a, b = ptr[j] - 1, ptr[j + 1] - 1
        row2[a:b] = row[a:b][rng.permutation(b - a)]
    s1, s2 = sp.SpLLT(nb=8), sp.SpLLT(nb=8)
    s1.analyse(n, ptr, row)
    s2.analyse(n, ptr, row2)
    assert np.array_equal(s1.order, s2.order) and np.array_equal(s1.blocks(), s2.blocks())


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_numeric_call_without_gpu_fails_loudly_not_fatally():
    """No CPU fallback: on a machine without a CUDA device every numeric entry point prints the reason
    and returns SPLLT_ERROR_UNKNOWN (-99) in info%flag -- it must neither compute anything nor abort
    the caller's process (the reference's error contract, src/spllt_data_mod.F90:31-35)."""
    import pytest
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    n, ptr, row, val = M.poisson2d(6)
    s = sp.SpLLT(nb=8)
    assert s.analyse(n, ptr, row) == 0           # host only
    assert s.factor(val) == -99                   # spllt_factor: needs the device
    x = np.ones(n)
    assert s.solve(x, 0) == -99 and np.all(x == 1.0)
    s.wait()                                      # must not crash either
    # multi-GPU bootstrap without a device: reported through the return code
    buf = (C.c_char * 128)()
    assert s.L.spllt_b200_comm_export(s.fkeep, buf) == -99
